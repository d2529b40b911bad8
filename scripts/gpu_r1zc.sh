#!/bin/bash
# Round-1 session zc: deeper host pipeline (6 slots, asynchronous expansion): parity + phase timings.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_drucker_prager.py tests/test_rust_models_adapters.py -m gpu -x -q > $OUT/pytest_r1zc.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r1zc.log
timeout 900 python scripts/host_wire_stats.py > $OUT/host_wire_stats_r1zc.jsonl 2> $OUT/host_wire_stats_r1zc.err; echo "stats rc=$?"; cut -c1-400 $OUT/host_wire_stats_r1zc.jsonl; tail -3 $OUT/host_wire_stats_r1zc.err
