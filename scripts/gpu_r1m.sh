#!/bin/bash
# Round-1 session m: full GPU suite after the FEM-kernel rewrite + host staging; timings; ncu of the new kernels.
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_r1m.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_r1m.log
timeout 600 python scripts/bench_newton.py --n 55 --steps 1 --newton-steps-only 20 --ab > $OUT/newton55_r1m.log 2>&1; echo "newton rc=$?"; tail -1 $OUT/newton55_r1m.log
timeout 600 python scripts/bench_models.py --steps 5 --out $OUT/models_r1m.json > $OUT/models_r1m.log 2>&1; echo "models rc=$?"; grep gather $OUT/models_r1m.log | tail -3
for mem in pageable pinned; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-memory $mem > $OUT/bench_${mem}_r1m.json 2> $OUT/bench_${mem}_r1m.err; echo "bench $mem rc=$?"
  python - <<PY
import json; d=json.loads(open("$OUT/bench_${mem}_r1m.json").read().strip().splitlines()[-1]); print("$mem", d["e2e"]["value"]/1e6, "MQP/s")
PY
done
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:"mises_form|qp_cell|gather_sum|gather_kernel" -s 30 -c 10 -o $OUT/prof_fem_r1m python scripts/bench_newton.py --n 55 --steps 1 --newton-steps-only 20 > $OUT/ncu_fem_r1m.log 2>&1; echo "ncu fem rc=$?"
ncu -i $OUT/prof_fem_r1m.ncu-rep --page raw --csv > $OUT/prof_fem_r1m_raw.csv 2>/dev/null
ncu -i $OUT/prof_fem_r1m.ncu-rep --page details > $OUT/prof_fem_r1m_details.txt 2>/dev/null
