#!/bin/bash
# round 2, GPU call 2: v4 kernels (persistent item fetch, bound-shifted softmax) through the GPU suite incl. the
# at-scale parity tests, kernel variants on the Reddit-shape probe, the bench line with its parity object, ncu.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/p2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p2_pytest.log
tail -15 gpurun_out/p2_pytest.log
CASE="reddit:232965:114615892:128:4"
for tag in "" px mb7 mb8 mb8u4 mb10u4; do
  echo "== variant '${tag}'" >> gpurun_out/p2_probe.log
  GTA_LIB_TAG=$tag timeout 300 python tools/agg_probe.py --cases $CASE --kinds gat gato spmm --col-blocks 3 --iters 10 >> gpurun_out/p2_probe.log 2>&1
done
GTA_LIB_TAG="" timeout 300 python tools/agg_probe.py --cases $CASE f64:232965:114615892:64:4 --kinds gat spmm --col-blocks 1 2 4 --iters 10 >> gpurun_out/p2_probe.log 2>&1
cat gpurun_out/p2_probe.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/p2_bench.json 2> gpurun_out/p2_bench.err
echo "bench rc=$?"; cat gpurun_out/p2_bench.json; tail -5 gpurun_out/p2_bench.err
timeout 300 python tools/agg_probe.py --cases $CASE --kinds gat --col-blocks 3 --iters 3 > gpurun_out/p2_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gat_aggregate -s 3 -c 1 -o gpurun_out/r02_gat_v4 \
  python tools/agg_probe.py --cases $CASE --kinds gat --col-blocks 3 --iters 3 > gpurun_out/p2_ncu.log 2>&1
echo "ncu rc=$?"
