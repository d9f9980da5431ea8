#!/bin/bash
# usage: scripts/bench_one.sh [bench args]  -- one-line summary of a 1-GPU bench run
python bench.py --no-cpu-baseline "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],4),'pcg/lm',d['lm']['pcg_iters_per_lm'],'launches',d['gpu_launches'])
"
