#!/bin/bash
# usage: scripts/profile_round.sh r1   (on the GPU box, under gpurun)
# plain run first (must exit 0), then the ncu launch list and one --set full capture of the main kernels;
# summarise afterwards with `python profiles/summarize.py <round>`.
set -e
R=${1:-r1}
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain_$R.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$R.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k 'regex:k_(eval|line_pass|cam_accum|cam_grad|pcg_cam|line_factor|cg_step|cg_begin|cam_finish|cam_step)' -c 36 \
    -f -o gpurun_out/prof_$R python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$R.log 2>&1
ls -la gpurun_out/prof_$R.ncu-rep gpurun_out/launches_$R.csv
