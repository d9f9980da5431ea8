#!/bin/bash
# usage: scripts/r2c_scale.sh N [config5]  -- one weak-scaling bench line at N GPUs (and the strong-scaling config-5 line)
cd /root/repo
N=$1
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
$T --master-port 29571 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_weak_n$N.log 2>&1; echo rc=$? >> gpurun_out/r2c_weak_n$N.log
if [ "$2" = "config5" ]; then
  $T --master-port 29572 bench.py --gpus $N --workload config5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_c5_n$N.log 2>&1; echo rc=$? >> gpurun_out/r2c_c5_n$N.log
fi
for f in gpurun_out/r2c_weak_n$N.log gpurun_out/r2c_c5_n$N.log; do
  [ -f $f ] || continue
  tail -2 $f | python -c "
import json,sys
L=sys.stdin.readlines()
print(L[-1].strip())
d=json.loads(L[-2])
p=d.get('parity') or {}
print('n_gpus',d['n_gpus'],d['config']['workload'][:40],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],4),'parity',p.get('ok'),p.get('per_trial_cost_max_rel_diff'))
print('class_avg_us',d['roofline']['class_avg_us'])
"
done
