#!/bin/bash
# round 2, GPU call 16 (1 GPU): next-item staging (GTA_ITEM_PREFETCH) and the publish without the extra fence, A/B on the
# Reddit-shape probe and the low-degree shape; then the GPU suite on the default build
set -u
mkdir -p gpurun_out
LOG=gpurun_out/p16_probe.log; : > $LOG
CASES="reddit:232965:114615892:128:4 lowdeg:232965:14326986:128:4"
for tag in "" nopf pffence; do
  echo "== variant '${tag:-default}'" >> $LOG
  GTA_LIB_TAG=$tag timeout 300 python tools/agg_probe.py --cases $CASES --kinds gat spmm --col-blocks 3 --chunk 1024 --iters 10 >> $LOG 2>&1
done
cat $LOG
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/p16_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p16_pytest.log
tail -4 gpurun_out/p16_pytest.log
bash tools/scale.sh 1 p16 --no-cpu-baseline
