#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python scripts/tune_variants.py --mises-only --lib fenics_constitutive_b200/libfcx_fmad.so > $OUT/tune_r1e_fmad.log 2>&1; echo "tune fmad rc=$?"; cat $OUT/tune_r1e_fmad.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ostage -s 3 -c 1 -f -o $OUT/prof_ostage_r1e \
  python bench.py --steps 1 --warmup 3 --e2e-steps 1 --e2e-qps 1000000 --no-cpu-baseline > $OUT/ncu_full_r1e.log 2>&1; echo "ncu full rc=$?"
ncu -i $OUT/prof_ostage_r1e.ncu-rep --page raw --csv > $OUT/prof_ostage_r1e_raw.csv 2>/dev/null
ncu -i $OUT/prof_ostage_r1e.ncu-rep --page details > $OUT/prof_ostage_r1e_details.txt 2>/dev/null
ls -la $OUT
