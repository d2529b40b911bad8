#!/bin/bash
# Round-1 session zl (N GPUs of one box, N = $1): torchrun paths of bench.py (native + reference arm) and of the
# Newton stand-in (weak scaling: every rank solves its own block, NCCL all-reduce of the scalars).
N=${1:-4}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_r1zl.txt; nproc >> $OUT/gpus_r1zl.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n${N}_r1zl.json 2> $OUT/bench_n${N}_r1zl.err; echo "bench n$N rc=$?"; tail -1 $OUT/bench_n${N}_r1zl.json; tail -3 $OUT/bench_n${N}_r1zl.err
timeout 600 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 3 > $OUT/bench_ref_n${N}_r1zl.json 2> $OUT/bench_ref_n${N}_r1zl.err; echo "bench ref n$N rc=$?"; tail -1 $OUT/bench_ref_n${N}_r1zl.json
timeout 900 $TR scripts/bench_newton.py --grid 55 --steps 2 --forcing ew > $OUT/newton55_n${N}_r1zl.log 2>&1; echo "newton n$N rc=$?"; tail -1 $OUT/newton55_n${N}_r1zl.log
