#!/bin/bash
# Round-2 session ag (1 GPU): pageable host path, ring slots x chunk size in combination (fewer / smaller slots keep
# more of the ring in the last-level cache; more chunks cost more CUDA calls on the enqueuing thread).
TAG=${1:-r2ag}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/e2e_sweep.py --kinds pageable --steps 3 --custom "3:32768;4:32768;3:16384;4:16384;3:65536;4:65536;2:32768" > $OUT/e2e_slots_$TAG.jsonl 2> $OUT/e2e_slots_$TAG.err; echo "rc=$?"
python - <<PY
import json
for l in open('$OUT/e2e_slots_$TAG.jsonl'):
    d=json.loads(l); print(d['tag'], d.get('slots','-'), d.get('chunk_qps','-'), d['MQPps_aggregate'], d['step_s'])
PY
tail -2 $OUT/e2e_slots_$TAG.err
