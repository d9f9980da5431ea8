#!/bin/bash
cd /root/repo
for sc in 1.0 0.5 0.25 0.1; do
  echo "== scale $sc"
  bash scripts/bench_shares.sh --scale $sc --cams 1000 2>&1 | tail -4
done | tee gpurun_out/scale_intercepts.log
