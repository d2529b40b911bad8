#!/bin/bash
# Round-2 session b (1 GPU): stress-only instantiation of the tile kernel (tests + sweep), e2e knob sweep
# (one process per memory kind: the pool only grows), fp64 DFMA peak, models block.
TAG=r2b
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest new"; timeout 1500 python -m pytest tests/test_gpu_round2.py tests/test_maps.py -m gpu -x -q > $OUT/pytest_new_$TAG.log 2>&1; echo "pytest new rc=$?"; tail -4 $OUT/pytest_new_$TAG.log
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_$TAG.log
echo "== dfma"; python -c "
from fenics_constitutive_b200._lib import lib
import json; print(json.dumps({'dfma_chain_peak_TFLOPs': lib().fcx_diag_dfma_peak()}))" | tee $OUT/dfma_peak_$TAG.json
echo "== stress-only sweep"; timeout 600 python scripts/tune_stress_only.py > $OUT/tune_stress_only_$TAG.jsonl 2>&1; echo "sweep rc=$?"; grep -E "occ|ctas_per_sm=(4|5|6|8)\"" $OUT/tune_stress_only_$TAG.jsonl
echo "== models"; timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 1 --e2e-qps 1000000 --e2e-memory pageable --no-cpu-baseline --no-newton > $OUT/bench_models_$TAG.json 2> $OUT/bench_models_$TAG.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_models_$TAG.json'))
for k,v in d['models'].items(): print(k, v.get('ms'), v.get('frac'))"
echo "== e2e sweep pageable"; timeout 900 python scripts/e2e_sweep.py --kinds pageable > $OUT/e2e_sweep_pageable_$TAG.jsonl 2>&1; echo "rc=$?"; cut -c1-220 $OUT/e2e_sweep_pageable_$TAG.jsonl
echo "== e2e sweep pinned"; timeout 900 python scripts/e2e_sweep.py --kinds pinned > $OUT/e2e_sweep_pinned_$TAG.jsonl 2>&1; echo "rc=$?"; cut -c1-220 $OUT/e2e_sweep_pinned_$TAG.jsonl
echo "== e2e sweep stress-only"; timeout 600 python scripts/e2e_sweep.py --quick --stress-only > $OUT/e2e_sweep_so_$TAG.jsonl 2>&1; echo "rc=$?"; cut -c1-220 $OUT/e2e_sweep_so_$TAG.jsonl
