#!/bin/bash
cd /root/repo
for v in "LBA_PCG_THREADS=128" "LBA_PCG_THREADS=64" "LBA_PCG_THREADS=32" "LBA_PCG_THREADS=256"; do
  echo "== cams 8000 $v"
  env $v bash scripts/bench_shares.sh --cams 8000 2>&1 | grep -v "^shares\|^per_kernel" | tail -2
done | tee gpurun_out/ab_pcg_threads.log
