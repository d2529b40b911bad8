#!/bin/bash
# Round-2 session s (1 GPU): mixed download wire (records + a share of the chunks by plain DMA) on page-locked arrays:
# bit-identity tests, then the share x pool-thread sweep.
TAG=${1:-r2s}
OUT=gpurun_out; mkdir -p $OUT
nproc > $OUT/host_$TAG.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $OUT/host_$TAG.txt
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host_path_memory_kinds" > $OUT/pytest_mix_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_mix_$TAG.log
echo "== sweep"; timeout 900 python scripts/e2e_sweep.py --kinds pinned --mix 0,15,25,35,50,100 --mix-threads 0,12 --steps 3 > $OUT/e2e_mix_$TAG.jsonl 2> $OUT/e2e_mix_$TAG.err; echo "rc=$?"
python - <<PY
import json
for l in open('$OUT/e2e_mix_$TAG.jsonl'):
    d=json.loads(l); print(d.get('threads',0), d['wire_mix'], d['wire_mix_used'], d['MQPps_aggregate'], d['step_s'])
PY
tail -2 $OUT/e2e_mix_$TAG.err
