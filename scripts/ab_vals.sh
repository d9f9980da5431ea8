#!/bin/bash
# usage: scripts/ab_vals.sh VAR v1 v2 ...   -- bench.py under VAR=v for each value ("-" = unset)
var=$1; shift
for v in "$@"; do
  echo "== $var=$v"
  if [ "$v" = "-" ]; then pre="env -u $var"; else pre="env $var=$v"; fi
  $pre python bench.py --no-cpu-baseline --steps 20 --warmup 3 | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],4))
"
done
