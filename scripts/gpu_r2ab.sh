#!/bin/bash
# Round-2 session ab (2 GPUs): look-ahead Krylov loop on a partitioned mesh: driver-equivalence tests (one GPU), partition
# check, config 5.
N=${1:-2}
TAG=${2:-r2ab}_n$N
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== driver tests (1 GPU)"; timeout 600 python -m pytest tests/test_solver_gpu.py -m gpu -q -k "krylov or cg or newton or readme or two_law" > $OUT/pytest_krylov_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_krylov_$TAG.log
echo "== check_partitioned_newton"
timeout 420 $TR scripts/check_partitioned_newton.py > $OUT/check_partitioned_$TAG.log 2>&1; echo "check rc=$?"
grep -E "degree|twin|ok|Error|error|assert" $OUT/check_partitioned_$TAG.log | cut -c1-330 | tail -9
echo "== bench_newton partition device"
timeout 240 $TR scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --partition --driver device > $OUT/newton55_part_device_$TAG.log 2>&1; echo "newton rc=$?"
tail -1 $OUT/newton55_part_device_$TAG.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('n_gpus','solve_s','linear_solve_s','ms_per_krylov_iteration','newton_iterations')}, [sum(k) for k in d['krylov_iterations']])"
