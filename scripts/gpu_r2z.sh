#!/bin/bash
# Round-2 final validation (1 GPU): smoke, every GPU test, bench both arms (all blocks), ncu launch list of the bench
# command, ncu full captures of the headline kernel and of the stress-only tile kernel.
TAG=${1:-r2z}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; lscpu | grep -E "Model name|^CPU\(s\)|NUMA|Socket" >> $OUT/gpu_$TAG.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_$TAG.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_reference_$TAG.json 2> $OUT/bench_reference_$TAG.err; echo "reference rc=$?"; cut -c1-300 $OUT/bench_reference_$TAG.json
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_$TAG.json'))
print('value', d['value'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])
e=d['e2e']; print('e2e pageable', e['value'], (e['roofline'] or {}).get('frac'), 'pinned', e['pinned']['value'], (e['pinned']['roofline'] or {}).get('frac'))
for k,v in d['models'].items(): print(k, v.get('ms'), v.get('frac'))
print(d['newton']); print(d['cpu_baseline'])"; tail -3 $OUT/bench_$TAG.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 3 --warmup 3 --e2e-steps 1 --e2e-qps 2000000 --e2e-memory pageable --no-cpu-baseline --no-newton > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (headline kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fcx_mises_ostage -s 3 -c 1 -f -o $OUT/prof_mises_$TAG \
  python bench.py --steps 1 --warmup 3 --e2e-steps 1 --e2e-qps 1000000 --e2e-memory pageable --no-cpu-baseline --no-newton --no-models > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
echo "== ncu full (stress-only tile kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fcx_tile_kernel -s 3 -c 1 -f -o $OUT/prof_mises_so_$TAG \
  python scripts/tune_stress_only.py --steps 2 > $OUT/ncu_full_so_$TAG.log 2>&1; echo "ncu full so rc=$?"
for f in prof_mises_$TAG prof_mises_so_$TAG; do
  ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/${f}_raw.csv 2>/dev/null
  ncu -i $OUT/$f.ncu-rep --page details > $OUT/${f}_details.txt 2>/dev/null
done
grep -E "Duration|DRAM Throughput|Registers Per|Achieved Occ|Issue Slots Busy" $OUT/prof_mises_so_${TAG}_details.txt
