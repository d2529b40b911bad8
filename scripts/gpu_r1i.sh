#!/bin/bash
# Round-1 session i: full GPU suite incl. solver / comfe-rs tests, bench, Newton stand-in bench.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu_r1i.txt 2>&1; nproc >> $OUT/gpu_r1i.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_r1i.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_r1i.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_r1i.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r1i.log; tail -25 $OUT/pytest_r1i.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_r1i.json 2> $OUT/bench_r1i.err; echo "bench rc=$?"; cat $OUT/bench_r1i.json; tail -5 $OUT/bench_r1i.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_r1i.json 2> $OUT/bench_ref_r1i.err; echo "bench ref rc=$?"; cat $OUT/bench_ref_r1i.json
timeout 900 python scripts/bench_newton.py --n 30 --steps 3 > $OUT/newton_r1i.log 2>&1; echo "newton rc=$?"; tail -30 $OUT/newton_r1i.log
