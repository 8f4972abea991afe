#!/bin/bash
# strong-scaling run of the headline bench on N GPUs of one box:  tools/scale.sh N [extra bench args]
N=$1; shift
mkdir -p gpurun_out
if [ "$N" = "1" ]; then timeout 400 python bench.py --gpus 1 "$@" > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; fi
tail -1 gpurun_out/scale_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=%d value=%.2f GTEPS ms=%.3f e2e=%.2f GTEPS (%.2f ms) kernels=%s alt=%s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms_by_name'], d['config'].get('alt_ms_per_step')))" || tail -20 gpurun_out/scale_n$N.err
