#!/bin/bash
# Round-2 session ae (1 GPU): compute-sanitizer on the kernels changed in this round (Drucker-Prager tile kernels with
# the fixed-(i, j) tangent store; Krylov loop with the parallel reduction tail, device-side stop, launch gate and
# alternating tile tickets): memcheck and racecheck.
TAG=${1:-r2ae}
OUT=gpurun_out; mkdir -p $OUT
export SANITIZE_ONLY=dp,krylov
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_small.py > $OUT/memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -4 $OUT/memcheck_$TAG.log
timeout 170 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_small.py > $OUT/racecheck_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -4 $OUT/racecheck_$TAG.log
