#!/bin/bash
# default bench line (GPU arm) and the reference arm, as the driver runs them
cd /root/repo
python bench.py > gpurun_out/bench_default_r2c.json 2> gpurun_out/bench_default_r2c.err; echo "rc=$?"
tail -c 600 gpurun_out/bench_default_r2c.json
python bench.py --impl reference > gpurun_out/bench_reference_r2c.json 2> gpurun_out/bench_reference_r2c.err; echo "rc=$?"
tail -c 400 gpurun_out/bench_reference_r2c.json
