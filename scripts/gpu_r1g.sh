#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_r1g.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_r1g.log
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > $OUT/racecheck_r1g.log 2>&1; echo "racecheck rc=$?"; tail -3 $OUT/racecheck_r1g.log
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > $OUT/memcheck_r1g.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck_r1g.log
timeout 900 python scripts/tune_variants.py > $OUT/tune_r1g.log 2>&1; echo "tune rc=$?"; cat $OUT/tune_r1g.log
