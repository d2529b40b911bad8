#!/bin/bash
# Round-1 session zs: last validation of the shipped state (smoke + every GPU test).
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_r1zs.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_r1zs.log
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_r1zs.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_r1zs.log
