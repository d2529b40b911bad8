#!/bin/bash
# Round-1 session u: Jacobian action from tangent records: parity, sanitizer, kernel timings, Newton solve.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_gather.py -m gpu -x -q > $OUT/pytest_r1u.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r1u.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > $OUT/racecheck_r1u.log 2>&1; echo "racecheck rc=$?"; tail -2 $OUT/racecheck_r1u.log
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > $OUT/memcheck_r1u.log 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/memcheck_r1u.log
timeout 900 python scripts/bench_newton.py --n 55 --steps 2 > $OUT/newton55_r1u.log 2>&1; echo "newton rc=$?"; tail -1 $OUT/newton55_r1u.log
