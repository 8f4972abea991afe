#!/bin/bash
# round 2, GPU call 1: chain-fold kernels through the whole GPU suite, kernel variants on the Reddit-shape probe,
# one ncu --set full capture of the new GAT kernel (after the same command exited 0 without ncu)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/p1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/p1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p1_pytest.log
tail -3 gpurun_out/p1_pytest.log
CASE="reddit:232965:114615892:128:4"
for tag in "" fx mb5 fx5 fx8; do
  echo "== variant '${tag}'" >> gpurun_out/p1_probe.log
  GTA_LIB_TAG=$tag timeout 300 python tools/agg_probe.py --cases $CASE --kinds gat spmm --col-blocks 3 1 --iters 10 >> gpurun_out/p1_probe.log 2>&1
done
cat gpurun_out/p1_probe.log
timeout 300 python tools/agg_probe.py --cases $CASE --kinds gat --col-blocks 3 --iters 3 > gpurun_out/p1_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gat_aggregate -s 3 -c 1 -o gpurun_out/r02_gat_v3 \
  python tools/agg_probe.py --cases $CASE --kinds gat --col-blocks 3 --iters 3 > gpurun_out/p1_ncu.log 2>&1
echo "ncu rc=$?"
