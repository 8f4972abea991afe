#!/bin/bash
# round 2, GPU call 3 (2 GPUs): GPU suite with the emulated fused exchange, bench at N=1, fused vs NCCL exchange at N=2
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/p3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p3_pytest.log
tail -25 gpurun_out/p3_pytest.log
CASE="reddit:232965:114615892:128:4"
for tag in "" nonc llh4; do
  echo "== variant '${tag}'" >> gpurun_out/p3_probe.log
  GTA_LIB_TAG=$tag timeout 300 python tools/agg_probe.py --cases $CASE --kinds gat spmm --col-blocks 3 --iters 10 >> gpurun_out/p3_probe.log 2>&1
done
cat gpurun_out/p3_probe.log
bash tools/scale.sh 1 p3
bash tools/scale.sh 2 p3fused --exchange fused
bash tools/scale.sh 2 p3nccl --exchange nccl
bash tools/scale.sh 2 p3gcn --exchange fused --workload reddit-gcn --no-e2e
