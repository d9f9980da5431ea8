#!/bin/bash
# usage: scripts/ab_variants.sh "A B C" [ENV=VAL ...]  -- bench each prebuilt library variant (amuya_sfm_lines_beta_b200/variants/liblba_X.so)
cd /root/repo
vars=$1; shift
cp amuya_sfm_lines_beta_b200/liblba.so /tmp/liblba_orig.so
for v in $vars; do
  cp amuya_sfm_lines_beta_b200/variants/liblba_$v.so amuya_sfm_lines_beta_b200/liblba.so
  echo "== variant $v $@"
  env "$@" python bench.py --no-cpu-baseline --steps 20 --warmup 3 | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1])
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),d['roofline']['class_avg_us'])
"
done
cp /tmp/liblba_orig.so amuya_sfm_lines_beta_b200/liblba.so
