#!/bin/bash
# Round-1 session za: staged gather kernel A/B; host pipeline phase timings.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gather.py tests/test_solver_gpu.py -m gpu -x -q > $OUT/pytest_r1za.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r1za.log
timeout 600 python scripts/bench_gather.py > $OUT/gather_ab_r1za.jsonl 2> $OUT/gather_ab_r1za.err; echo "gather rc=$?"; cat $OUT/gather_ab_r1za.jsonl; tail -3 $OUT/gather_ab_r1za.err
timeout 600 python scripts/host_wire_stats.py > $OUT/host_wire_stats_r1za.jsonl 2> $OUT/host_wire_stats_r1za.err; echo "stats rc=$?"; cat $OUT/host_wire_stats_r1za.jsonl; tail -3 $OUT/host_wire_stats_r1za.err
