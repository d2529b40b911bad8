#!/bin/bash
# Round-2 session u (1 GPU): per-region sampling of the fused form() kernel.
TAG=${1:-r2u}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fcx_mises_form_kernel -s 4 -c 1 -f -o $OUT/prof_form_$TAG \
  python scripts/bench_newton.py --grid 55 --steps 1 --forcing ew --driver device --newton-steps-only 20 > $OUT/ncu_form_$TAG.log 2>&1; echo "ncu form rc=$?"
for f in prof_form_$TAG; do
  ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/${f}_raw.csv 2>/dev/null
  ncu -i $OUT/$f.ncu-rep --page details > $OUT/${f}_details.txt 2>/dev/null
  ncu -i $OUT/$f.ncu-rep --page source --csv > $OUT/${f}_source.csv 2>/dev/null
  grep -E "Duration|DRAM Throughput|Registers Per|Achieved Occ|Issue Slots Busy|Theoretical Occ|Executed Ipc" $OUT/${f}_details.txt
done
python scripts/ncu_regions.py $OUT/prof_form_${TAG}_source.csv 60
