#!/bin/bash
# Round-2 session af (1 GPU): whole GPU test suite on the final state (new finite-difference tangent tests included).
TAG=${1:-r2af}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_$TAG.log
