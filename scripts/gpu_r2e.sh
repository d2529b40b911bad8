#!/bin/bash
# Round-2 session e (N GPUs): where the wall time of the partitioned Newton solve goes (device vs python Krylov driver).
N=${1:-2}
TAG=r2e_n$N
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for drv in device python; do
  timeout 240 $TR scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --partition --driver $drv > $OUT/newton55_part_${drv}_$TAG.log 2>&1; echo "newton partition $drv rc=$?"
  tail -1 $OUT/newton55_part_${drv}_$TAG.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('n_gpus','cg_driver','solve_s','linear_solve_s','residual_s','update_s','krylov_setup_s','ms_per_krylov_iteration','setup_s','newton_iterations')})"
done
