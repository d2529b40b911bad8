#!/bin/bash
# Round-2 session p (1 GPU): staged gather kernel after the instruction diet (node-granular cp.async staging, 16-byte
# table loads and result stores): parity tests, timing, ncu.
TAG=${1:-r2p}
OUT=gpurun_out; mkdir -p $OUT
echo "== tests"; timeout 900 python -m pytest tests/test_gather.py tests/test_gpu_round2.py tests/test_solver_gpu.py -m gpu -q -k "gather or form or fused or two_law or readme" > $OUT/pytest_gather_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gather_$TAG.log
echo "== timing"; timeout 600 python scripts/bench_gather.py --variants 1,0 --ctas 0 --reps 20 > $OUT/gather_ab_$TAG.jsonl 2> $OUT/gather_ab_$TAG.err; echo "rc=$?"; cut -c1-300 $OUT/gather_ab_$TAG.jsonl; tail -2 $OUT/gather_ab_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_staged_kernel -s 6 -c 1 -f -o $OUT/prof_gather_$TAG \
  python scripts/bench_gather.py --variants 1 --ctas 0 --reps 5 > $OUT/ncu_gather_$TAG.log 2>&1; echo "ncu gather rc=$?"
for f in prof_gather_$TAG; do
  ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/${f}_raw.csv 2>/dev/null
  ncu -i $OUT/$f.ncu-rep --page details > $OUT/${f}_details.txt 2>/dev/null
  ncu -i $OUT/$f.ncu-rep --page source --csv > $OUT/${f}_source.csv 2>/dev/null
  grep -E "Duration|DRAM Throughput|Registers Per|Achieved Occ|Issue Slots Busy|Theoretical Occ|Executed Ipc" $OUT/${f}_details.txt
done
