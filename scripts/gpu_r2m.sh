#!/bin/bash
# Round-2 session m (1 GPU): ncu full capture of the Drucker-Prager tile kernel (classic, VAR 1).
TAG=${1:-r2m}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fcx_tile_kernel -s 4 -c 1 -f -o $OUT/prof_dp_$TAG \
  python scripts/tune_dp.py --qps 4000000 --variants 1 --steps 2 > $OUT/ncu_dp_$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/prof_dp_$TAG.ncu-rep --page raw --csv > $OUT/prof_dp_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_dp_$TAG.ncu-rep --page details > $OUT/prof_dp_${TAG}_details.txt 2>/dev/null
ncu -i $OUT/prof_dp_$TAG.ncu-rep --page source --csv > $OUT/prof_dp_${TAG}_source.csv 2>/dev/null
grep -E "Duration|DRAM Throughput|Registers Per|Achieved Occ|Issue Slots Busy|Theoretical Occ|Executed Ipc|No Eligible" $OUT/prof_dp_${TAG}_details.txt
