#!/bin/bash
# round 2, GPU call 11 (8 GPUs): the headline and RMAT-24 with the adaptive item size
set -u
mkdir -p gpurun_out
bash tools/scale.sh 8 p11fused --exchange fused
bash tools/scale.sh 8 p11rmat24 --workload rmat24-gcn --exchange fused --steps 5 --warmup 3
bash tools/scale.sh 4 p11fused --exchange fused --no-e2e
