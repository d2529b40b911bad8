#!/bin/bash
# Round-2 session k (1 GPU): CTA-parallel tail of the Krylov reductions -- solver tests, config-5 solve, and the
# per-kernel launch list of the Krylov loop.
TAG=${1:-r2k}
OUT=gpurun_out; mkdir -p $OUT
echo "== solver tests"; timeout 600 python -m pytest tests/test_solver_gpu.py tests/test_gpu_round2.py -m gpu -q -x -k "krylov or cg or newton or solver or two_law or readme" > $OUT/pytest_solver_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_solver_$TAG.log
for drv in device python; do
  echo "== bench_newton $drv"
  timeout 240 python scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --driver $drv > $OUT/newton55_${drv}_$TAG.log 2>&1; echo "newton rc=$?"
  tail -1 $OUT/newton55_${drv}_$TAG.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('n_gpus','cg_driver','solve_s','linear_solve_s','residual_s','ms_per_krylov_iteration','setup_s','newton_iterations','kernel_ms')})"
done
echo "== ncu launch list of the Krylov loop"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gsum|cg_update|qp_cell|halo|kr_" -s 300 -c 90 --csv --log-file $OUT/krylov_launches_$TAG.csv \
  python scripts/bench_newton.py --grid 55 --steps 1 --forcing ew --driver device > $OUT/ncu_krylov_$TAG.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv,collections
rows=[r for r in csv.reader(l for l in open('$OUT/krylov_launches_$TAG.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
acc=collections.defaultdict(list)
for r in rows[1:]:
    acc[r[ki][:60]].append(float(r[vi].replace(',','')))
for k,v in acc.items(): print(k, len(v), 'launches, mean', sum(v)/len(v), rows[1][ui])
PY
