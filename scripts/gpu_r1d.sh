#!/bin/bash
# r1d: parity of the new kernels, sanitizer, variant sweep
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu_r1d.txt 2>&1; nproc >> $OUT/gpu_r1d.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_r1d.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r1d.log; tail -5 $OUT/pytest_r1d.log
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > $OUT/memcheck_r1d.log 2>&1; echo "memcheck rc=$?"; tail -4 $OUT/memcheck_r1d.log
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > $OUT/racecheck_r1d.log 2>&1; echo "racecheck rc=$?"; tail -4 $OUT/racecheck_r1d.log
timeout 900 python scripts/tune_variants.py > $OUT/tune_r1d.log 2>&1; echo "tune rc=$?"; cat $OUT/tune_r1d.log
