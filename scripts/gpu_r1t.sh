#!/bin/bash
# Round-1 session t: final-state evidence: full GPU suite, bench (+reference arm), launch list, ncu --set full of the
# Mises / FEM / Drucker-Prager kernels, all-model timings, Newton stand-in.
OUT=gpurun_out; mkdir -p $OUT; T=r1t
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu_$T.txt 2>&1; nproc >> $OUT/gpu_$T.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$T.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_$T.log
timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest_$T.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$T.log; tail -6 $OUT/pytest_$T.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_$T.json 2> $OUT/bench_$T.err; echo "bench rc=$?"; cat $OUT/bench_$T.json; tail -3 $OUT/bench_$T.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $OUT/bench_ref_$T.json 2> $OUT/bench_ref_$T.err; echo "bench ref rc=$?"; cat $OUT/bench_ref_$T.json
timeout 900 python scripts/bench_models.py --steps 10 --out $OUT/models_$T.json > $OUT/models_$T.log 2>&1; echo "models rc=$?"; grep -E "rs_|gather" $OUT/models_$T.log
timeout 900 python scripts/bench_newton.py --n 55 --steps 2 > $OUT/newton55_$T.log 2>&1; echo "newton rc=$?"; tail -1 $OUT/newton55_$T.log
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_$T.csv \
  python bench.py --steps 3 --warmup 3 --e2e-steps 1 --e2e-qps 1000000 --no-cpu-baseline > $OUT/ncu_launch_$T.log 2>&1; echo "ncu list rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:fcx_mises_ostage -s 3 -c 1 -o $OUT/prof_mises_$T python bench.py --steps 1 --warmup 3 --e2e-steps 1 --e2e-qps 1000000 --no-cpu-baseline > $OUT/ncu_mises_$T.log 2>&1; echo "ncu mises rc=$?"
timeout 900 $NCU -k regex:"mises_form|qp_cell|gather_sum" -s 30 -c 6 -o $OUT/prof_fem_$T python scripts/bench_newton.py --n 55 --steps 1 --newton-steps-only 20 > $OUT/ncu_fem_$T.log 2>&1; echo "ncu fem rc=$?"
timeout 900 $NCU --kernel-name-base demangled -k regex:"DruckerPrager|MisesLin|gather_kernel" -c 8 -o $OUT/prof_rs_$T python scripts/bench_models.py --qps 4000000 --steps 1 > $OUT/ncu_rs_$T.log 2>&1; echo "ncu rs rc=$?"
for f in mises fem rs; do
  ncu -i $OUT/prof_${f}_$T.ncu-rep --page raw --csv > $OUT/prof_${f}_${T}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_${f}_$T.ncu-rep --page details > $OUT/prof_${f}_${T}_details.txt 2>/dev/null
done
ls -la $OUT | grep $T
