#!/bin/bash
# usage: scripts/r2c_mgpu.sh N -- GPU tests (multi-GPU ones included) and the weak-scaling bench line at N GPUs
cd /root/repo
N=$1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_n$N.log
tail -3 gpurun_out/pytest_gpu_n$N.log
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29571 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_weak_n$N.log 2>&1; echo rc=$? >> gpurun_out/r2c_weak_n$N.log
tail -2 gpurun_out/r2c_weak_n$N.log | python -c "
import json,sys
L=sys.stdin.readlines()
print(L[-1].strip())
d=json.loads(L[-2])
print('n_gpus',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],4),'parity',d.get('parity'))
print('class_avg_us',d['roofline']['class_avg_us'])
"
