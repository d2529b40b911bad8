#!/bin/bash
# Round-2 session t (N GPUs, N = 4 or 8): partition check (ranks with two neighbours) and config 5 on ONE mesh over the
# ranks after the overlapped ghost push and the parallel reduction tail.
N=${1:-4}
TAG=r2t_n$N
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== check_partitioned_newton"
timeout 420 $TR scripts/check_partitioned_newton.py > $OUT/check_partitioned_$TAG.log 2>&1; echo "check rc=$?"
grep -E "degree|twin|ok|Error|error|assert" $OUT/check_partitioned_$TAG.log | cut -c1-300 | tail -9
echo "== bench_newton partition device"
timeout 240 $TR scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --partition --driver device > $OUT/newton55_part_device_$TAG.log 2>&1; echo "newton rc=$?"
tail -1 $OUT/newton55_part_device_$TAG.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('n_gpus','cg_driver','solve_s','linear_solve_s','residual_s','ms_per_krylov_iteration','setup_s','newton_iterations')})"
