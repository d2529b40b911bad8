#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -k "mises" > $OUT/pytest_r1f.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_r1f.log
timeout 600 python scripts/tune_variants.py --mises-only > $OUT/tune_r1f.log 2>&1; echo "tune rc=$?"; cat $OUT/tune_r1f.log
timeout 600 python scripts/tune_variants.py --mises-only --lib fenics_constitutive_b200/libfcx_fmad.so > $OUT/tune_r1f_fmad.log 2>&1; echo "tune fmad rc=$?"; cat $OUT/tune_r1f_fmad.log
