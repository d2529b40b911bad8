#!/bin/bash
# Round-1 session n: shuffle reduce-scatter cell kernel + 2-load node gathers: parity, sanitizer, timing, ncu.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_gather.py -m gpu -x -q > $OUT/pytest_r1n.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_r1n.log
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > $OUT/memcheck_r1n.log 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/memcheck_r1n.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > $OUT/racecheck_r1n.log 2>&1; echo "racecheck rc=$?"; tail -2 $OUT/racecheck_r1n.log
timeout 600 python scripts/bench_newton.py --n 55 --steps 1 --newton-steps-only 20 --ab > $OUT/newton55_r1n.log 2>&1; echo "newton rc=$?"; tail -1 $OUT/newton55_r1n.log
timeout 600 python scripts/bench_models.py --steps 5 --out $OUT/models_r1n.json > $OUT/models_r1n.log 2>&1; echo "models rc=$?"; grep gather $OUT/models_r1n.log | tail -3
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:"mises_form|qp_cell|gather_sum" -s 30 -c 8 -o $OUT/prof_fem_r1n python scripts/bench_newton.py --n 55 --steps 1 --newton-steps-only 20 > $OUT/ncu_fem_r1n.log 2>&1; echo "ncu fem rc=$?"
timeout 600 $NCU -k regex:gather_kernel -c 2 -o $OUT/prof_gather_r1n python scripts/bench_models.py --qps 2000000 --steps 2 > $OUT/ncu_gather_r1n.log 2>&1; echo "ncu gather rc=$?"
for f in fem gather; do
ncu -i $OUT/prof_${f}_r1n.ncu-rep --page raw --csv > $OUT/prof_${f}_r1n_raw.csv 2>/dev/null
ncu -i $OUT/prof_${f}_r1n.ncu-rep --page details > $OUT/prof_${f}_r1n_details.txt 2>/dev/null
done
