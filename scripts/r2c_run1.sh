#!/bin/bash
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
bash scripts/bench_shares.sh 2>&1 | tail -5 | tee gpurun_out/shares_c1.log
