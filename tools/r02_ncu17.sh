#!/bin/bash
# round 2, GPU call 17 (1 GPU): where does a GAT item's fixed cost go?  --set full of gat_aggregate_kernel on the low-degree
# shape (61 edges per row, 20 per item: the per-item work dominates), after a plain run of the same command exited 0.
set -u
mkdir -p gpurun_out
P="tools/agg_probe.py --cases lowdeg:232965:14326986:128:4 --kinds gat --col-blocks 3 --chunk 1024 --iters 1"
timeout 300 python $P > gpurun_out/p17_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gat_aggregate_kernel" -s 3 -c 1 -o gpurun_out/r02_gat_lowdeg python $P > gpurun_out/p17_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/p17_plain.log; ls -la gpurun_out/r02_gat_lowdeg.ncu-rep
