#!/bin/bash
# round 2, GPU call 5 (8 GPUs): the headline at N=8 with the in-kernel exchange, the NCCL baseline beside it,
# N=4, and BASELINE config 5 (RMAT-24 GCN on 8 GPUs)
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/p5_topo.txt 2>&1
bash tools/scale.sh 8 p5fused --exchange fused
bash tools/scale.sh 8 p5nccl --exchange nccl --no-e2e
bash tools/scale.sh 4 p5fused --exchange fused --no-e2e
bash tools/scale.sh 8 p5rmat24 --workload rmat24-gcn --exchange fused --steps 5 --warmup 3
