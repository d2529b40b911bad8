#!/bin/bash
# Round-2 session aa (1 GPU): device-side residual test + look-ahead block enqueue of the Krylov loop: solver tests,
# config 5 with and without look-ahead.
TAG=${1:-r2aa}
OUT=gpurun_out; mkdir -p $OUT
echo "== solver tests"; timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_gpu_round2.py tests/test_mesh_partition.py -m gpu -q -k "krylov or cg or newton or solver or two_law or readme or partition or mises or plastic" > $OUT/pytest_solver_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_solver_$TAG.log
for la in 1 0; do
  echo "== bench_newton lookahead=$la"
  FCX_KRYLOV_LOOKAHEAD=$la timeout 240 python scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --driver device > $OUT/newton55_la${la}_$TAG.log 2>&1; echo "newton rc=$?"
  tail -1 $OUT/newton55_la${la}_$TAG.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('solve_s','linear_solve_s','residual_s','ms_per_krylov_iteration','newton_iterations','mean_sigma_xx')}, [sum(k) for k in d['krylov_iterations']])"
done
