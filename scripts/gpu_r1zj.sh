#!/bin/bash
# Round-1 session zj: full validation of the final state: smoke, all GPU tests, bench (both arms), ncu launch list +
# full capture of the headline kernel, per-model table, Newton stand-in, sanitizer.
bash scripts/gpu_check.sh r1zj
OUT=gpurun_out
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $OUT/bench_reference_r1zj.json 2> $OUT/bench_reference_r1zj.err; echo "reference rc=$?"; cat $OUT/bench_reference_r1zj.json
timeout 900 python scripts/bench_newton.py --n 55 --steps 2 > $OUT/newton55_r1zj.log 2>&1; echo "newton rc=$?"; tail -1 $OUT/newton55_r1zj.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > $OUT/racecheck_r1zj.log 2>&1; echo "racecheck rc=$?"; tail -2 $OUT/racecheck_r1zj.log
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > $OUT/memcheck_r1zj.log 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/memcheck_r1zj.log
timeout 300 python bench.py --steps 5 --warmup 3 --e2e-memory pageable --no-cpu-baseline > $OUT/bench_pageable_r1zj.json 2>/dev/null; echo "pageable rc=$?"; python -c "import json;d=json.load(open('$OUT/bench_pageable_r1zj.json'));print(d['e2e'])"
