#!/bin/bash
# round 2, GPU call 12 (1 GPU): relaxed polling + even item split; chunk sweep down to 128; bf16 exchange emulation
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/p12_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p12_pytest.log
tail -8 gpurun_out/p12_pytest.log
CASE="reddit:232965:114615892:128:4"
for ch in 128 256 512 1024; do
  timeout 300 python tools/agg_probe.py --cases $CASE --kinds gat spmm --col-blocks 3 --chunk $ch --iters 10 >> gpurun_out/p12_probe.log 2>&1
done
# 1/8 of the Reddit shape on one GPU (what a rank of 8 sees, minus the exchange): item-size effect on the tail
for ch in 128 256 1024; do
  timeout 300 python tools/agg_probe.py --cases eighth:232965:14326986:128:4 --kinds gat --col-blocks 3 --chunk $ch --iters 20 >> gpurun_out/p12_probe.log 2>&1
done
cat gpurun_out/p12_probe.log
bash tools/scale.sh 1 p12
