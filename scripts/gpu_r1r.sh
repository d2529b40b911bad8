#!/bin/bash
# Round-1 session r: packed download wire of the Mises host path: parity, e2e for pinned / pageable arrays, with and without.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_r1r.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_r1r.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_r1r.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_r1r.log
for mem in pinned pageable; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-memory $mem --e2e-steps 3 > $OUT/bench_${mem}_r1r.json 2> $OUT/bench_${mem}_r1r.err; echo "bench $mem rc=$?"; tail -2 $OUT/bench_${mem}_r1r.err
  python - <<PY
import json; d=json.loads(open("$OUT/bench_${mem}_r1r.json").read().strip().splitlines()[-1]); print("$mem", d["e2e"]["value"]/1e6, "MQP/s")
PY
done
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_host.py > $OUT/memcheck_host_r1r.log 2>&1; echo "memcheck host rc=$?"; tail -2 $OUT/memcheck_host_r1r.log
