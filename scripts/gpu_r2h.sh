#!/bin/bash
# Round-2 session h (1 GPU): warp-uniform-point gather kernel (variant 3): tests, A/B, ncu.
TAG=r2h
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gather"; timeout 900 python -m pytest tests/test_gather.py -m gpu -x -q > $OUT/pytest_gather_$TAG.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_gather_$TAG.log
echo "== gather A/B"; timeout 600 python scripts/bench_gather.py --variants 1,3 --ctas 0,4,6,8 > $OUT/gather_ab_$TAG.jsonl 2>&1; echo "rc=$?"; cut -c1-200 $OUT/gather_ab_$TAG.jsonl
echo "== ncu gather wq"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_wq -s 2 -c 1 -f -o $OUT/prof_gather_$TAG \
  python scripts/bench_gather.py --reps 3 --ctas 0 --variants 3 > $OUT/ncu_gather_$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/prof_gather_$TAG.ncu-rep --page raw --csv > $OUT/prof_gather_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_gather_$TAG.ncu-rep --page details > $OUT/prof_gather_${TAG}_details.txt 2>/dev/null
grep -E "Duration|DRAM Throughput|L1/TEX Cache Throughput|Registers Per|Achieved Occ|Issue Slots Busy|Executed Ipc Active" $OUT/prof_gather_${TAG}_details.txt
