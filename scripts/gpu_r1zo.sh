#!/bin/bash
# Round-1 session zo: host path (wire kernels, all modes) under compute-sanitizer.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_host.py > $OUT/memcheck_host_r1zo.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck_host_r1zo.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_host.py > $OUT/racecheck_host_r1zo.log 2>&1; echo "racecheck rc=$?"; tail -3 $OUT/racecheck_host_r1zo.log
timeout 900 compute-sanitizer --tool initcheck python scripts/sanitize_host.py > $OUT/initcheck_host_r1zo.log 2>&1; echo "initcheck rc=$?"; tail -3 $OUT/initcheck_host_r1zo.log
