#!/bin/bash
# Round-2 session d (N GPUs, default 1): device-resident Krylov loop (fcx_krylov.cu): tests, A/B against the Python /
# NCCL driver on the 1 M-cell Newton stand-in; with N > 1 the partitioned check and the partitioned solve.
N=${1:-1}
TAG=r2d_n$N
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [ "$N" = "1" ]; then
  echo "== pytest solver"; timeout 1500 python -m pytest tests/test_solver_gpu.py tests/test_gpu_round2.py -m gpu -x -q > $OUT/pytest_solver_$TAG.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_solver_$TAG.log
  for drv in device python; do
    timeout 600 python scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --driver $drv > $OUT/newton55_${drv}_$TAG.log 2>&1; echo "newton $drv rc=$?"; tail -1 $OUT/newton55_${drv}_$TAG.log | cut -c1-1300
  done
  timeout 600 python scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --driver device --check-every 50 > $OUT/newton55_device_ce50_$TAG.log 2>&1; echo "rc=$?"; tail -1 $OUT/newton55_device_ce50_$TAG.log | cut -c1-600
else
  timeout 240 $TR scripts/check_partitioned_newton.py > $OUT/check_partitioned_$TAG.log 2>&1; rc=$?; echo "check rc=$rc"; grep -v "^\*\|OMP_NUM\|^$" $OUT/check_partitioned_$TAG.log | tail -14 | cut -c1-330
  if [ "$rc" != "0" ]; then echo "partitioned check failed: stopping"; exit 1; fi
  for drv in device python; do
    timeout 240 $TR scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --partition --driver $drv > $OUT/newton55_part_${drv}_$TAG.log 2>&1; echo "newton partition $drv rc=$?"; tail -1 $OUT/newton55_part_${drv}_$TAG.log | cut -c1-1300
  done
  timeout 420 $TR bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_$TAG.json'))
e=d['e2e']; print('value', d['value'], 'e2e pageable', e['value'], 'wire', e['wire'], 'pinned', e['pinned']['value'], 'wire', e['pinned']['wire'])
print(d['newton'])"; tail -3 $OUT/bench_$TAG.err
  timeout 240 $TR bench.py --gpus $N --impl reference --steps 5 --warmup 3 > $OUT/bench_reference_$TAG.json 2>/dev/null; cat $OUT/bench_reference_$TAG.json | cut -c1-400
fi
