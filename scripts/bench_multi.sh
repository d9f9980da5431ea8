#!/bin/bash
# usage: scripts/bench_multi.sh N [extra bench args]  -- torchrun bench on N GPUs, print value + class table
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1])
print('n_gpus',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],4), 'pcg/lm', d['lm']['pcg_iters_per_lm'])
print('class_avg_us',d['roofline']['class_avg_us'])
print('shares',d['roofline']['time_shares'])
"
