#!/bin/bash
# Round-2 session r (N GPUs): config 5 (ONE mesh over the ranks) and the bench line under torchrun after the
# parallel reduction tail.
N=${1:-2}
TAG=r2r_n$N
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== bench_newton partition device"
timeout 240 $TR scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --partition --driver device > $OUT/newton55_part_device_$TAG.log 2>&1; echo "newton rc=$?"
tail -1 $OUT/newton55_part_device_$TAG.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('n_gpus','cg_driver','solve_s','linear_solve_s','residual_s','ms_per_krylov_iteration','setup_s','newton_iterations')})"
echo "== bench.py --gpus $N"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$OUT/bench_$TAG.json'))
print('value', d['value'], 'frac', d['roofline']['frac'], 'n_gpus', d['n_gpus'], d['clocks'])
e=d['e2e']; print('e2e pageable', e['value'], 'pinned', e['pinned']['value'])
print(d['newton'])"; tail -3 $OUT/bench_$TAG.err
