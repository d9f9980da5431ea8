#!/bin/bash
# bench.py (no CPU baseline) -> value, e2e and the per-class device-time table
python bench.py --no-cpu-baseline --steps 20 --warmup 3 "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],4))
print('shares',d['roofline']['time_shares'])
print('class_avg_us',d['roofline']['class_avg_us'])
print('per_kernel',{k:(v['avg_us'],v['frac']) for k,v in d['roofline']['per_kernel'].items()})
"
