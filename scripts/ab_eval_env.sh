#!/bin/bash
# usage: scripts/ab_eval_env.sh "ENV1=.. ENV2=.." "ENV.." ...  -- Jacobian-kernel time (stand-alone, 30 launches) under each environment
cd /root/repo
for v in "$@"; do
  echo "== $v"
  env $v python tests/measure_eval_kernel.py 30 2>&1 | tail -1
done | tee gpurun_out/ab_eval_env.log
