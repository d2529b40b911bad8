#!/bin/bash
# Round-2 session a (1 GPU): smoke, all GPU tests (incl. stress-only, 16 M strided samples, map kernels, cell-list
# form), bench both arms with the new blocks (models, e2e pageable + pinned + host roofline, newton),
# stress-only kernel sweep, ncu launch list + full captures (headline kernel, stress-only kernel).
TAG=r2a
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; lscpu | grep -E "Model name|^CPU\(s\)|NUMA|Socket" >> $OUT/gpu_$TAG.txt; free -g | head -2 >> $OUT/gpu_$TAG.txt
nvidia-smi topo -m >> $OUT/gpu_$TAG.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
echo "== pytest new"; timeout 1500 python -m pytest tests/test_gpu_round2.py tests/test_maps.py -m gpu -x -q > $OUT/pytest_new_$TAG.log 2>&1; echo "pytest new rc=$?"; tail -12 $OUT/pytest_new_$TAG.log
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest_$TAG.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cut -c1-3000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_reference_$TAG.json 2> $OUT/bench_reference_$TAG.err; echo "reference rc=$?"; cat $OUT/bench_reference_$TAG.json; tail -3 $OUT/bench_reference_$TAG.err
echo "== stress-only sweep"; timeout 600 python scripts/tune_stress_only.py > $OUT/tune_stress_only_$TAG.jsonl 2>&1; echo "sweep rc=$?"; cat $OUT/tune_stress_only_$TAG.jsonl
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 3 --warmup 3 --e2e-steps 1 --e2e-qps 2000000 --e2e-memory pageable --no-cpu-baseline --no-newton > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (headline kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fcx_mises_ostage -s 3 -c 1 -f -o $OUT/prof_mises_$TAG \
  python bench.py --steps 1 --warmup 3 --e2e-steps 1 --e2e-qps 1000000 --e2e-memory pageable --no-cpu-baseline --no-newton --no-models > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
echo "== ncu full (stress-only kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fcx_mises_ostage -s 3 -c 1 -f -o $OUT/prof_mises_so_$TAG \
  python scripts/tune_stress_only.py --steps 2 > $OUT/ncu_full_so_$TAG.log 2>&1; echo "ncu full so rc=$?"
for f in prof_mises_$TAG prof_mises_so_$TAG; do
  ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/${f}_raw.csv 2>/dev/null
  ncu -i $OUT/$f.ncu-rep --page details > $OUT/${f}_details.txt 2>/dev/null
done
ls -la $OUT | head -50
