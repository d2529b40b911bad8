#!/bin/bash
# Round-2 session f (N GPUs on one box): bench.py both arms (headline, e2e pageable + pinned with the auto wire,
# newton = config 5 partitioned over the N ranks), e2e knob mini-sweep with the ranks sharing the host.
N=${1:-4}
TAG=r2f_n$N
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
nproc > $OUT/host_$TAG.txt; lscpu | grep -E "Model name|^CPU\(s\)|NUMA|Socket" >> $OUT/host_$TAG.txt; free -g | head -2 >> $OUT/host_$TAG.txt; nvidia-smi topo -m >> $OUT/host_$TAG.txt 2>&1
timeout 300 $TR bench.py --gpus $N --impl reference --steps 5 --warmup 3 2>/dev/null | grep "^{" > $OUT/bench_reference_$TAG.json; echo "reference rc=$?"; cut -c1-330 $OUT/bench_reference_$TAG.json
timeout 480 $TR bench.py --gpus $N --steps 10 --warmup 3 2> $OUT/bench_$TAG.err | grep "^{" > $OUT/bench_$TAG.json; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_$TAG.json'))
e=d['e2e']; print('value', d['value']/1e9, 'e2e pageable', e['value']/1e6, 'wire', e['wire'], 'frac', (e['roofline'] or {}).get('frac'), '| pinned', e['pinned']['value']/1e6, 'wire', e['pinned']['wire'], 'frac', (e['pinned']['roofline'] or {}).get('frac'))
print(e['host_probes'])
print(d['newton'])"; tail -3 $OUT/bench_$TAG.err | cut -c1-300
timeout 400 $TR scripts/e2e_sweep.py --mini --kinds pinned 2>/dev/null | grep "^{" > $OUT/e2e_sweep_pinned_$TAG.jsonl; echo "sweep pinned rc=$?"; cut -c1-200 $OUT/e2e_sweep_pinned_$TAG.jsonl
timeout 400 $TR scripts/e2e_sweep.py --mini --kinds pageable 2>/dev/null | grep "^{" > $OUT/e2e_sweep_pageable_$TAG.jsonl; echo "sweep pageable rc=$?"; cut -c1-200 $OUT/e2e_sweep_pageable_$TAG.jsonl
