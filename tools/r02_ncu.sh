#!/bin/bash
# round 2 ncu evidence (1 GPU): launch list of the Reddit-shape GAT step, --set full of the dominant kernel and the
# GEMM, --set full of the weighted aggregate (Reddit-shape GCN).  Every ncu command follows a plain run of the SAME
# command line that exited 0.
set -u
mkdir -p gpurun_out
A="bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --no-graph"
B="bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --no-graph --workload reddit-gcn"
timeout 300 python $A > gpurun_out/ncu_plain_a.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_reddit_gat.csv python $A > gpurun_out/ncu_a1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gat_aggregate_kernel|gemm_tc_kernel" -s 12 -c 2 -o gpurun_out/r02_gat_layer_final python $A > gpurun_out/ncu_a2.log 2>&1
echo "A rc=$?"
timeout 300 python $B > gpurun_out/ncu_plain_b.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:aggregate_kernel" -s 6 -c 1 -o gpurun_out/r02_gcn_aggregate_final python $B > gpurun_out/ncu_b.log 2>&1
echo "B rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_reddit_gat.csv
bash tools/scale.sh 1 p10
