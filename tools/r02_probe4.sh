#!/bin/bash
# round 2, GPU call 4 (1 GPU): suite after batch grabs / publish rewrite, aggregate occupancy variants, other workloads
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/p4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p4_pytest.log
tail -5 gpurun_out/p4_pytest.log
CASE="reddit:232965:114615892:128:4"
for tag in "" t8 t9; do
  echo "== variant '${tag}'" >> gpurun_out/p4_probe.log
  GTA_LIB_TAG=$tag timeout 300 python tools/agg_probe.py --cases $CASE f64:232965:114615892:64:4 --kinds gat spmm --col-blocks 3 --iters 10 >> gpurun_out/p4_probe.log 2>&1
done
cat gpurun_out/p4_probe.log
bash tools/scale.sh 1 p4
bash tools/scale.sh 1 p4heavy --workload reddit-heavy-gat --no-cpu-baseline
bash tools/scale.sh 1 p4gcn --workload reddit-gcn --no-cpu-baseline
bash tools/scale.sh 1 p4rmat20 --workload rmat20-gcn --no-cpu-baseline
bash tools/scale.sh 1 p4flickr --workload flickr-gcn --no-cpu-baseline
bash tools/scale.sh 1 p4cora --workload cora-gat --no-cpu-baseline
bash tools/scale.sh 1 p4corah8 --workload cora-gat --heads 8 --no-cpu-baseline
