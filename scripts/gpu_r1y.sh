#!/bin/bash
# Round-1 session y: gather kernel with two QPs per thread: parity (incl. fused == unfused bit for bit), timing, ncu.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_gather.py -m gpu -x -q > $OUT/pytest_r1y.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r1y.log
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > $OUT/racecheck_r1y.log 2>&1; echo "racecheck rc=$?"; tail -2 $OUT/racecheck_r1y.log
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > $OUT/memcheck_r1y.log 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/memcheck_r1y.log
timeout 900 python scripts/bench_models.py --steps 5 --out $OUT/models_r1y.json > $OUT/models_r1y.log 2>&1; echo "models rc=$?"; grep -E "gather" $OUT/models_r1y.log | tail -3
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:gather_kernel -c 2 -o $OUT/prof_gather_r1y python scripts/bench_models.py --qps 2000000 --steps 2 > $OUT/ncu_gather_r1y.log 2>&1; echo "ncu gather rc=$?"
ncu -i $OUT/prof_gather_r1y.ncu-rep --page raw --csv > $OUT/prof_gather_r1y_raw.csv 2>/dev/null
