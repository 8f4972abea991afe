#!/bin/bash
# round 2, GPU call 18 (1 GPU): bound cache + fast block division + single-exp chain merge + cheap exchange gate, against
# the stale 'nopf' build (= the kernel of call 14); then the GPU suite and the headline line on the default build
set -u
mkdir -p gpurun_out
LOG=gpurun_out/p18_probe.log; : > $LOG
CASES="reddit:232965:114615892:128:4 lowdeg:232965:14326986:128:4"
for tag in "" nopf; do
  echo "== variant '${tag:-default}'" >> $LOG
  GTA_LIB_TAG=$tag timeout 300 python tools/agg_probe.py --cases $CASES --kinds gat gatb spmm --col-blocks 3 --chunk 1024 --iters 10 >> $LOG 2>&1
done
cat $LOG
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/p18_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p18_pytest.log
tail -4 gpurun_out/p18_pytest.log
bash tools/scale.sh 1 p18 --no-cpu-baseline
