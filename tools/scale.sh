#!/bin/bash
# strong-scaling run of the headline bench on N GPUs of one box:  tools/scale.sh N TAG [extra bench args]
N=$1; TAG=$2; shift; shift
mkdir -p gpurun_out
OUT=gpurun_out/scale_${TAG}_n$N
if [ "$N" = "1" ]; then timeout 300 python bench.py --gpus 1 "$@" > $OUT.json 2> $OUT.err
else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" > $OUT.json 2> $OUT.err; fi
echo "rc=$?"
tail -1 $OUT.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d.get('e2e') or {}; p=d.get('parity') or {}
print('N=%d value=%.2f GTEPS ms=%.3f e2e=%.2f GTEPS (%.2f ms, h2d alone %.2f ms) kernels=%s parity=%.3f bitwise=%s enqueue=%.3f ms' % (d['n_gpus'], d['value'], d['ms_per_step'], e.get('value', 0), e.get('ms_per_step', 0), e.get('h2d_alone_ms', 0), {k: round(v, 4) for k, v in d['roofline']['kernel_ms_by_name'].items()}, p.get('max_err_over_tol', -1), p.get('bitwise_rerun'), d['config']['host_enqueue_ms_per_step']))" || tail -30 $OUT.err
