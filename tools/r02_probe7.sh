#!/bin/bash
# round 2, GPU call 7 (1 GPU): suite; static striding on RMAT; bf16 with 32 lanes per item; bf16 bench lines
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/p7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p7_pytest.log
tail -15 gpurun_out/p7_pytest.log
CASE="reddit:232965:114615892:128:4"
timeout 300 python tools/agg_probe.py --cases $CASE --kinds gat gatb spmm spmmb --col-blocks 3 1 --iters 10 > gpurun_out/p7_probe.log 2>&1
timeout 300 python tools/agg_probe.py --cases rmat20:0:0:256:4 rmat20:0:0:128:4 rmat22:0:0:256:4 --kinds spmm spmmb --col-blocks 1 --iters 5 >> gpurun_out/p7_probe.log 2>&1
cat gpurun_out/p7_probe.log
bash tools/scale.sh 1 p7bf16 --dtype bf16 --no-cpu-baseline
bash tools/scale.sh 1 p7gcnbf16 --dtype bf16 --workload reddit-gcn --no-cpu-baseline --no-e2e
bash tools/scale.sh 1 p7rmat20 --workload rmat20-gcn --no-cpu-baseline --no-e2e
bash tools/scale.sh 1 p7flickr --workload flickr-gcn --no-cpu-baseline
