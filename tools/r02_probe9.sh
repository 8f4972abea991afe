#!/bin/bash
# round 2, GPU call 9 (1 GPU): adaptive chunk; chunk sweep at N=1; two-layer Flickr GCN; headline
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/p9_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p9_pytest.log
tail -8 gpurun_out/p9_pytest.log
CASE="reddit:232965:114615892:128:4"
for ch in 256 512 1024; do
  timeout 300 python tools/agg_probe.py --cases $CASE --kinds gat spmm --col-blocks 3 --chunk $ch --iters 10 >> gpurun_out/p9_probe.log 2>&1
done
cat gpurun_out/p9_probe.log
bash tools/scale.sh 1 p9
bash tools/scale.sh 1 p9flickr2 --workload flickr-gcn2 --no-cpu-baseline
bash tools/scale.sh 1 p9heavy --workload reddit-heavy-gat --no-cpu-baseline --no-e2e
