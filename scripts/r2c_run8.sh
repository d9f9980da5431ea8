#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
bash scripts/bench_shares.sh 2>&1 | grep -v "^shares\|^per_kernel" | tail -2 | tee gpurun_out/shares_tab.log
bash scripts/bench_shares.sh --cams 8000 2>&1 | grep -v "^shares\|^per_kernel" | tail -2 | tee -a gpurun_out/shares_tab.log
