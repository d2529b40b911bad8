#!/bin/bash
# Round-2 session ah (2 GPUs): the bench line under torchrun on the final state.
N=${1:-2}
TAG=r2ah_n$N
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 2 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('$OUT/bench_$TAG.json').read().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'n_gpus', d['n_gpus'], d['clocks'])
e=d['e2e']; print('e2e pageable', e['value'], 'pinned', e['pinned']['value'])
print(d['newton'])
print({k:(v['ms'],v['frac']) for k,v in d['models'].items() if k.startswith('rs_')})"; tail -2 $OUT/bench_$TAG.err
