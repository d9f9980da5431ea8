#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for v in "LBA_LP_TAB=0" "LBA_LP_TAB=1"; do
  echo "== $v"
  env $v bash scripts/bench_shares.sh 2>&1 | grep -v "^shares\|^per_kernel" | tail -2
done | tee gpurun_out/ab_lptab.log
