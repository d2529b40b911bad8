#!/bin/bash
# Round-1 session l: QP-parallel FEM kernels (gather / residual / Jacobian action): parity, sanitizer, timing A/B.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_gather.py -m gpu -x -q > $OUT/pytest_r1l.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_r1l.log
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > $OUT/memcheck_r1l.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck_r1l.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > $OUT/racecheck_r1l.log 2>&1; echo "racecheck rc=$?"; tail -3 $OUT/racecheck_r1l.log
timeout 600 python scripts/bench_newton.py --n 55 --steps 1 --newton-steps-only 20 --ab > $OUT/newton55_r1l.log 2>&1; echo "newton rc=$?"; tail -2 $OUT/newton55_r1l.log
timeout 600 python scripts/bench_models.py --steps 5 --out $OUT/models_r1l.json > $OUT/models_r1l.log 2>&1; echo "models rc=$?"; grep gather $OUT/models_r1l.log | tail -3
