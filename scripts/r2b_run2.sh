#!/bin/bash
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for v in "LBA_EVAL_DC=1" "LBA_EVAL_DC=0" ; do
  echo "== $v"
  env $v bash scripts/bench_shares.sh 2>&1 | tail -5
done > gpurun_out/ab_eval.log 2>&1
cat gpurun_out/ab_eval.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on \
    -k 'regex:k_eval_dc' -c 2 \
    -f -o gpurun_out/prof_b python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_b.log 2>&1
