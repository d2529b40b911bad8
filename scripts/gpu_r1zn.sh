#!/bin/bash
# N ranks sharing one host: e2e with the pool divided among the ranks (default) vs 8 threads per rank.
N=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --e2e-steps 3 > $OUT/bench_n${N}_r1zn.json 2> $OUT/bench_n${N}_r1zn.err; echo "bench n$N rc=$?"; tail -1 $OUT/bench_n${N}_r1zn.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['e2e']['value']/1e6)"
