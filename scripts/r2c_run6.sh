#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for cams in 1000 2000 8000; do for z in 0 1; do echo "CAMS=$cams LBA_EVAL_Z=$z"; CAMS=$cams LBA_EVAL_Z=$z python tests/measure_eval_kernel.py 30 2>&1 | tail -1; done; done | tee gpurun_out/ab_z_cams.log
