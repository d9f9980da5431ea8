N=$1
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
$T --master-port 29571 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_weak_n$N.log 2>&1; echo rc=$? >> gpurun_out/r2_weak_n$N.log
$T --master-port 29572 bench.py --gpus $N --workload config5 --steps 20 --warmup 3 > gpurun_out/r2_c5_n$N.log 2>&1; echo rc=$? >> gpurun_out/r2_c5_n$N.log
