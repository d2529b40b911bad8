#!/bin/bash
# Round-2 session x (1 GPU): why the newton block of bench.py is slower per Krylov iteration than scripts/bench_newton.py.
TAG=${1:-r2x}
OUT=gpurun_out; mkdir -p $OUT
timeout 240 python scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --driver device > $OUT/newton55_device_$TAG.log 2>&1; echo "newton rc=$?"
tail -1 $OUT/newton55_device_$TAG.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench_newton', {k:d[k] for k in ('solve_s','linear_solve_s','ms_per_krylov_iteration')})"
for flags in "--no-models --e2e-memory pageable" "--e2e-memory pageable"; do
  timeout 600 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline $flags > $OUT/bench_x_$TAG.json 2> $OUT/bench_x_$TAG.err; echo "bench rc=$? ($flags)"
  python -c "
import json; d=json.loads(open('$OUT/bench_x_$TAG.json').read().splitlines()[-1]); n=d['newton']; print('bench.py', {k:n[k] for k in ('solve_s','linear_solve_s','ms_per_krylov_iteration')})"
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu,clocks_throttle_reasons.active --format=csv
