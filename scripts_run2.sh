set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py --workload cora-gat --steps 5 --warmup 3 2>&1 | tail -5
python bench.py --workload flickr-gcn --steps 5 --warmup 3 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_reddit.log 2>&1; tail -5 gpurun_out/bench_reddit.log
