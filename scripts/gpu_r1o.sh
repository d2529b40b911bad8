#!/bin/bash
# Round-1 session o: Drucker-Prager kernels, quad-shuffle nodal gathers, s_next race fix: full suite, sanitizer, timings.
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_r1o.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_r1o.log
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > $OUT/memcheck_r1o.log 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/memcheck_r1o.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > $OUT/racecheck_r1o.log 2>&1; echo "racecheck rc=$?"; tail -4 $OUT/racecheck_r1o.log
timeout 600 python scripts/bench_newton.py --n 55 --steps 1 --newton-steps-only 20 > $OUT/newton55_r1o.log 2>&1; echo "newton rc=$?"; tail -1 $OUT/newton55_r1o.log
timeout 900 python scripts/bench_models.py --steps 5 --out $OUT/models_r1o.json > $OUT/models_r1o.log 2>&1; echo "models rc=$?"; grep -E "gather|rs_" $OUT/models_r1o.log | tail -6
