#!/bin/bash
# ncu evidence for the headline workload (run on the GPU box, after a plain run exited 0).
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_reddit_gat.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gat_aggregate_kernel -s 3 -c 2 -f -o gpurun_out/prof_gat_aggregate $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
