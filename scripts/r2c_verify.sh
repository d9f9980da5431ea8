#!/bin/bash
# final check on N GPUs: smoke, GPU tests (multi-GPU ones when N >= 2), the bench line at N
cd /root/repo
N=${1:-1}
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final_n$N.log
tail -3 gpurun_out/pytest_gpu_final_n$N.log
if [ "$N" = "1" ]; then
  python bench.py --no-cpu-baseline > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "bench rc=$?"
  tail -1 gpurun_out/bench_final_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],4),'frac',r['frac'],'frac_survey',r['frac_survey_bytes'],'launches',d['gpu_launches'])
print(r['class_avg_us'])"
else
  bash scripts/r2c_scale.sh $N
fi
