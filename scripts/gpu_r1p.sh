#!/bin/bash
# Round-1 session p: full suite after reverting the quad-shuffle gathers; kernel timings; DP ncu.
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_r1p.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_r1p.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > $OUT/racecheck_r1p.log 2>&1; echo "racecheck rc=$?"; tail -3 $OUT/racecheck_r1p.log
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > $OUT/memcheck_r1p.log 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/memcheck_r1p.log
timeout 600 python scripts/bench_newton.py --n 55 --steps 1 --newton-steps-only 20 > $OUT/newton55k_r1p.log 2>&1; echo "newton rc=$?"; tail -1 $OUT/newton55k_r1p.log
timeout 900 python scripts/bench_models.py --steps 5 --out $OUT/models_r1p.json > $OUT/models_r1p.log 2>&1; echo "models rc=$?"; grep -E "gather|rs_" $OUT/models_r1p.log | tail -6
timeout 900 python scripts/bench_newton.py --n 55 --steps 2 > $OUT/newton55_r1p.log 2>&1; echo "newton full rc=$?"; tail -1 $OUT/newton55_r1p.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:"DruckerPrager" -s 6 -c 2 -o $OUT/prof_dp_r1p python scripts/bench_models.py --qps 4000000 --steps 1 > $OUT/ncu_dp_r1p.log 2>&1; echo "ncu dp rc=$?"
ncu -i $OUT/prof_dp_r1p.ncu-rep --page raw --csv > $OUT/prof_dp_r1p_raw.csv 2>/dev/null
ncu -i $OUT/prof_dp_r1p.ncu-rep --page details > $OUT/prof_dp_r1p_details.txt 2>/dev/null
