#!/bin/bash
# Round-2 session j (2 GPUs): validation of the overlapped Krylov loop (ghost push without waiting, interior cells
# first, one-thread wait kernel, boundary cells) -- full partition check (device + python drivers, twin of the
# reference's test_mpi_solver) and the config-5 solve on ONE mesh over the ranks.
N=${1:-2}
TAG=r2j_n$N
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== check_partitioned_newton"
timeout 420 $TR scripts/check_partitioned_newton.py > $OUT/check_partitioned_$TAG.log 2>&1; echo "check rc=$?"
grep -E "degree|twin|ok|Error|error|assert" $OUT/check_partitioned_$TAG.log | cut -c1-400 | tail -12
for rep in 1 2; do
  echo "== bench_newton partition device (rep $rep)"
  timeout 240 $TR scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --partition --driver device > $OUT/newton55_part_device_${TAG}_$rep.log 2>&1; echo "newton rc=$?"
  tail -1 $OUT/newton55_part_device_${TAG}_$rep.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('n_gpus','cg_driver','solve_s','linear_solve_s','residual_s','ms_per_krylov_iteration','setup_s','newton_iterations')})"
done
