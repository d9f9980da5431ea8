#!/bin/bash
cd /root/repo
for z in 0 1; do echo "LBA_EVAL_Z=$z"; LBA_EVAL_Z=$z python tests/measure_eval_kernel.py 30 2>&1 | tail -1; done | tee gpurun_out/ab_z.log
LBA_LIB=amuya_sfm_lines_beta_b200/variants/liblba_clock.so python tests/measure_eval_kernel.py 1 > gpurun_out/clock.log 2>&1; grep "^Z" gpurun_out/clock.log | tail -6
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
bash scripts/bench_shares.sh 2>&1 | tail -4 | tee gpurun_out/shares_z.log
