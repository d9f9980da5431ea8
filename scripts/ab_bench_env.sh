#!/bin/bash
# usage: scripts/ab_bench_env.sh "ENV1=.." "ENV2=.." ...  -- bench value and per-class times under each environment
cd /root/repo
for v in "$@"; do
  echo "== $v"
  env $v bash scripts/bench_shares.sh 2>&1 | grep "^value\|^class"
done | tee gpurun_out/ab_bench_env.log
