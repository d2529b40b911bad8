#!/bin/bash
# Round-1 session q (2 GPUs): torchrun paths of bench.py (native + reference arm) and of the Newton stand-in.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_r1q.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_n2_r1q.json 2> $OUT/bench_n2_r1q.err; echo "bench n2 rc=$?"; tail -1 $OUT/bench_n2_r1q.json; tail -3 $OUT/bench_n2_r1q.err
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > $OUT/bench_ref_n2_r1q.json 2> $OUT/bench_ref_n2_r1q.err; echo "bench ref n2 rc=$?"; tail -1 $OUT/bench_ref_n2_r1q.json
timeout 900 $TR scripts/bench_newton.py --n 55 --steps 1 > $OUT/newton55_n2_r1q.log 2>&1; echo "newton n2 rc=$?"; tail -1 $OUT/newton55_n2_r1q.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/bench_n1_r1q.json 2> $OUT/bench_n1_r1q.err; echo "bench n1 rc=$?"; tail -1 $OUT/bench_n1_r1q.json
