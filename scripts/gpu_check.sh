#!/bin/bash
# One GPU session: smoke, GPU parity tests, bench, ncu launch list + full capture.
# Usage (from the repo root, under gpurun): bash scripts/gpu_check.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $OUT/gpu_$TAG.txt
echo "== smoke" | tee $OUT/smoke_$TAG.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke_$TAG.log
tail -3 $OUT/smoke_$TAG.log
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -15 $OUT/pytest_$TAG.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 3 --warmup 3 --e2e-steps 1 --e2e-qps 2000000 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fcx_mises_ostage -s 3 -c 1 -f -o $OUT/prof_mises_$TAG \
  python bench.py --steps 1 --warmup 3 --e2e-steps 1 --e2e-qps 1000000 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
ncu -i $OUT/prof_mises_$TAG.ncu-rep --page raw --csv > $OUT/prof_mises_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_mises_$TAG.ncu-rep --page details > $OUT/prof_mises_${TAG}_details.txt 2>/dev/null
timeout 600 python scripts/bench_models.py --out $OUT/models_$TAG.json > $OUT/models_$TAG.log 2>&1; echo "bench_models rc=$?"; tail -3 $OUT/models_$TAG.log
