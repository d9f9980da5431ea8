#!/bin/bash
# usage: scripts/tune_cam_threads.sh "32 64 128" [bench args] -- sweep the CTA size of the per-camera block-sum kernels
sizes=$1; shift
for t in $sizes; do
  echo "== LBA_CAM_THREADS=$t"
  LBA_CAM_THREADS=$t bash scripts/bench_shares.sh "$@" | grep -v "^shares\|^per_kernel"
done
