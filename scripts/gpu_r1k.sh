#!/bin/bash
# Round-1 session k: pageable/registered e2e, Newton stand-in at the ~1M-cell config,
# ncu --set full of the gather / fused form / Jacobian-action / elastic / Kelvin kernels.
OUT=gpurun_out; mkdir -p $OUT
for mem in pageable registered; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-memory $mem > $OUT/bench_${mem}_r1k.json 2> $OUT/bench_${mem}_r1k.err; echo "bench $mem rc=$?"
  python - <<PY
import json; d=json.loads(open("$OUT/bench_${mem}_r1k.json").read().strip().splitlines()[-1]); print("$mem", d["e2e"])
PY
done
timeout 900 python scripts/bench_newton.py --n 55 --steps 2 > $OUT/newton55_r1k.log 2>&1; echo "newton55 rc=$?"; tail -3 $OUT/newton55_r1k.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:gather_kernel -c 2 -o $OUT/prof_gather_r1k python scripts/bench_models.py --qps 2000000 --steps 2 > $OUT/ncu_gather_r1k.log 2>&1; echo "ncu gather rc=$?"
timeout 900 $NCU -k regex:"mises_form|cell_kernel|gather_sum" -s 20 -c 8 -o $OUT/prof_newton_r1k python scripts/bench_newton.py --n 55 --steps 1 > $OUT/ncu_newton_r1k.log 2>&1; echo "ncu newton rc=$?"
timeout 600 $NCU -k regex:"tile_kernel|uniaxial" -c 12 -o $OUT/prof_models_r1k python scripts/bench_models.py --steps 1 > $OUT/ncu_models_r1k.log 2>&1; echo "ncu models rc=$?"
for f in gather newton models; do
  ncu -i $OUT/prof_${f}_r1k.ncu-rep --page raw --csv > $OUT/prof_${f}_r1k_raw.csv 2>/dev/null
  ncu -i $OUT/prof_${f}_r1k.ncu-rep --page details > $OUT/prof_${f}_r1k_details.txt 2>/dev/null
done
ls -la $OUT | tail -20
