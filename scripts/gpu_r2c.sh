#!/bin/bash
# Round-2 session c (1 GPU): thread-per-cell gather kernel (tests, A/B over CTAs per SM, ncu), DP record fix,
# pageable defaults of the host pipeline, full test suite.
TAG=r2c
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gather"; timeout 900 python -m pytest tests/test_gather.py tests/test_drucker_prager.py -m gpu -x -q > $OUT/pytest_gather_$TAG.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_gather_$TAG.log
echo "== gather A/B"; timeout 600 python scripts/bench_gather.py > $OUT/gather_ab_$TAG.jsonl 2>&1; echo "rc=$?"; cut -c1-200 $OUT/gather_ab_$TAG.jsonl
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_$TAG.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --e2e-memory pageable,pinned,registered > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_$TAG.json'))
e=d['e2e']; print('e2e pageable', e['value'], e['roofline']['frac'], 'pinned', e['pinned']['value'], 'registered', e.get('registered',{}).get('value'), e.get('registered',{}).get('register_s'))
for k,v in d['models'].items(): print(k, v.get('ms'), v.get('frac'))
print(d['newton'])"; tail -3 $OUT/bench_$TAG.err
echo "== ncu gather"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_cell -s 2 -c 1 -f -o $OUT/prof_gather_$TAG \
  python scripts/bench_gather.py --reps 3 --ctas 0 > $OUT/ncu_gather_$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/prof_gather_$TAG.ncu-rep --page raw --csv > $OUT/prof_gather_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_gather_$TAG.ncu-rep --page details > $OUT/prof_gather_${TAG}_details.txt 2>/dev/null
grep -E "Duration|DRAM Throughput|L1/TEX Cache Throughput|Registers Per|Achieved Occ|Issue Slots Busy|Executed Ipc Active" $OUT/prof_gather_${TAG}_details.txt
