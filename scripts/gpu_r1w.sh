#!/bin/bash
# Round-1 session w: node-major element-vector slots: solver tests, sanitizer, Newton solve, ncu of the new J_apply pair.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_gather.py -m gpu -x -q > $OUT/pytest_r1w.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r1w.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > $OUT/racecheck_r1w.log 2>&1; echo "racecheck rc=$?"; tail -2 $OUT/racecheck_r1w.log
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > $OUT/memcheck_r1w.log 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/memcheck_r1w.log
timeout 900 python scripts/bench_newton.py --n 55 --steps 2 > $OUT/newton55_r1w.log 2>&1; echo "newton rc=$?"; tail -1 $OUT/newton55_r1w.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:"qp_cell|gather_sum|pcg_" -s 40 -c 10 -o $OUT/prof_fem_r1w python scripts/bench_newton.py --n 55 --steps 1 --newton-steps-only 20 > $OUT/ncu_fem_r1w.log 2>&1; echo "ncu fem rc=$?"
ncu -i $OUT/prof_fem_r1w.ncu-rep --page raw --csv > $OUT/prof_fem_r1w_raw.csv 2>/dev/null
