#!/bin/bash
# Round-2 session ad (8 GPUs): config 5 on ONE mesh over 8 ranks with the final Krylov loop.
N=${1:-8}
TAG=r2ad_n$N
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --partition --driver device > $OUT/newton55_part_device_$TAG.log 2>&1; echo "newton rc=$?"
tail -1 $OUT/newton55_part_device_$TAG.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('n_gpus','solve_s','linear_solve_s','ms_per_krylov_iteration','newton_iterations','setup_s')}, [sum(k) for k in d['krylov_iterations']])"
tail -3 $OUT/newton55_part_device_$TAG.log | cut -c1-300
