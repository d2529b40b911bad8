#!/bin/bash
# Round-1 session zt: GPU twins of the remaining reference solver tests (cyclic Mises, creep under traction, plane strain
# vs 3D, Kelvin vs Maxwell) + the whole GPU suite once more.
OUT=gpurun_out; mkdir -p $OUT
timeout 700 python -m pytest tests -m gpu -q > $OUT/pytest_r1zt.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_r1zt.log
