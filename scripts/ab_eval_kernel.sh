#!/bin/bash
# usage: scripts/ab_eval_kernel.sh "name1 name2 ..."  -- Jacobian-kernel time of each prebuilt variant library
cd /root/repo
for v in $1; do
  LBA_LIB=amuya_sfm_lines_beta_b200/variants/liblba_$v.so python tests/measure_eval_kernel.py 30 2>&1 | tail -1
done | tee gpurun_out/ab_eval_kernel.log
