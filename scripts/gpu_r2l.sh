#!/bin/bash
# Round-2 session l (1 GPU): Drucker-Prager kernels with the slow fp64 operations spelled for latency (VAR 1)
# against the reference's spelling (VAR 0): GPU parity tests, A/B timing, ncu capture of the classic kernel.
TAG=${1:-r2l}
OUT=gpurun_out; mkdir -p $OUT
echo "== tests"; timeout 900 python -m pytest tests/test_drucker_prager.py tests/test_rust_models_adapters.py tests/test_gpu_round2.py tests/test_solver_gpu.py -m gpu -q -k "drucker or Drucker or rust or Rust or device_krylov or stress_only" > $OUT/pytest_dp_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_dp_$TAG.log
echo "== A/B"; timeout 600 python scripts/tune_dp.py --variants 0,1 --ctas 0 > $OUT/tune_dp_$TAG.jsonl 2> $OUT/tune_dp_$TAG.err; echo "tune rc=$?"; cat $OUT/tune_dp_$TAG.jsonl | cut -c1-260; tail -2 $OUT/tune_dp_$TAG.err
echo "== ncu full (classic, VAR 1)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fcx_tile_kernel -s 4 -c 1 -f -o $OUT/prof_dp_$TAG \
  python scripts/tune_dp.py --qps 4000000 --variants 1 --steps 2 > $OUT/ncu_dp_$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/prof_dp_$TAG.ncu-rep --page raw --csv > $OUT/prof_dp_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_dp_$TAG.ncu-rep --page details > $OUT/prof_dp_${TAG}_details.txt 2>/dev/null
ncu -i $OUT/prof_dp_$TAG.ncu-rep --page source --csv > $OUT/prof_dp_${TAG}_source.csv 2>/dev/null
grep -E "Duration|DRAM Throughput|Registers Per|Achieved Occ|Issue Slots Busy|Theoretical Occ|Executed Ipc|No Eligible" $OUT/prof_dp_${TAG}_details.txt
