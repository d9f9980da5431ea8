#!/bin/bash
# usage: scripts/bench_trace.sh N [bench args] -- N-GPU bench with the device-side exchange trace (LBA_PEER_TRACE)
N=$1; shift
LBA_PEER_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --warmup 3 --no-cpu-baseline --no-profile "$@" > gpurun_out/trace$N.log 2>&1
grep "peer trace" gpurun_out/trace$N.log | sort
grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"pcg_iters_per_lm": [0-9.]*' gpurun_out/trace$N.log | head -4 | tr '\n' ' '; echo
