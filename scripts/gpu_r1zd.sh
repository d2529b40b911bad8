#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_drucker_prager.py tests/test_rust_models_adapters.py -m gpu -x -q > $OUT/pytest_r1zf.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r1zf.log
FCX_TIMELINE_SKIPS=1 timeout 600 python scripts/host_timeline.py 8000000 $OUT > $OUT/host_timeline_r1zf.jsonl 2> $OUT/host_timeline_r1zf.err; echo "rc=$?"; cut -c1-330 $OUT/host_timeline_r1zf.jsonl; tail -3 $OUT/host_timeline_r1zf.err
timeout 600 python scripts/host_wire_stats.py > $OUT/host_wire_stats_r1zf.jsonl 2> $OUT/host_wire_stats_r1zf.err; echo "stats rc=$?"; cut -c1-120 $OUT/host_wire_stats_r1zf.jsonl; tail -3 $OUT/host_wire_stats_r1zf.err
