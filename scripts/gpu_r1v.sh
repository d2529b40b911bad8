#!/bin/bash
# Round-1 session v: fused PCG vector kernels + tangent records: solver tests, Newton solve timing.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_solver_gpu.py -m gpu -x -q > $OUT/pytest_r1v.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r1v.log
timeout 900 python scripts/bench_newton.py --n 55 --steps 2 > $OUT/newton55_r1v.log 2>&1; echo "newton rc=$?"; tail -1 $OUT/newton55_r1v.log
timeout 600 compute-sanitizer --tool racecheck python scripts/bench_newton.py --n 6 --steps 1 > $OUT/racecheck_newton_r1v.log 2>&1; echo "racecheck rc=$?"; tail -2 $OUT/racecheck_newton_r1v.log
