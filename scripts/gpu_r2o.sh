#!/bin/bash
# Round-2 session o (1 GPU): ncu full capture (with the SASS-level sampling page) of the staged gather kernel and of
# the Krylov element kernel.
TAG=${1:-r2o}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_staged_kernel -s 6 -c 1 -f -o $OUT/prof_gather_$TAG \
  python scripts/bench_gather.py --variants 1 --ctas 0 --reps 5 > $OUT/ncu_gather_$TAG.log 2>&1; echo "ncu gather rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qp_cell_kernel -s 60 -c 1 -f -o $OUT/prof_qpcell_$TAG \
  python scripts/bench_newton.py --grid 55 --steps 1 --forcing ew --driver device --newton-steps-only 100 > $OUT/ncu_qpcell_$TAG.log 2>&1; echo "ncu qp_cell rc=$?"
for f in prof_gather_$TAG prof_qpcell_$TAG; do
  ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/${f}_raw.csv 2>/dev/null
  ncu -i $OUT/$f.ncu-rep --page details > $OUT/${f}_details.txt 2>/dev/null
  ncu -i $OUT/$f.ncu-rep --page source --csv > $OUT/${f}_source.csv 2>/dev/null
  grep -E "Duration|DRAM Throughput|Registers Per|Achieved Occ|Issue Slots Busy|Theoretical Occ|Executed Ipc" $OUT/${f}_details.txt
done
