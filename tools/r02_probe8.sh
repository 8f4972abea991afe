#!/bin/bash
# round 2, GPU call 8 (8 GPUs): grouped slot blocks + cheap chain publish; copy-CTA sweep; N=4; RMAT-24
set -u
mkdir -p gpurun_out
bash tools/scale.sh 8 p8fused --exchange fused
bash tools/scale.sh 8 p8fused_c192 --exchange fused --copy-ctas 192 --no-e2e
bash tools/scale.sh 8 p8fused_c48 --exchange fused --copy-ctas 48 --no-e2e
bash tools/scale.sh 4 p8fused --exchange fused --no-e2e
bash tools/scale.sh 8 p8rmat24 --workload rmat24-gcn --exchange fused --steps 5 --warmup 3
bash tools/scale.sh 8 p8heavy --workload reddit-heavy-gat --exchange fused --no-e2e
