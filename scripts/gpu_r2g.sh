#!/bin/bash
# Round-2 session g (1 GPU): two-phase Drucker-Prager update (tests, throughput, ncu), nodal-increment pre-pass of the
# gather, relaxed-flag Krylov loop; full test suite.
TAG=r2g
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_$TAG.log
echo "== models"; timeout 600 python scripts/bench_models.py --out $OUT/models_$TAG.json > $OUT/models_$TAG.log 2>&1; echo "rc=$?"; python -c "
import json
for r in json.load(open('$OUT/models_$TAG.json')):
    print(r['kernel'], round(r['ms'],4), round(r.get('frac_of_measured_hbm', 0),3), r.get('plastic_fraction'))"
echo "== ncu DP"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:DruckerPrager -s 3 -c 1 -f -o $OUT/prof_dp_$TAG \
  python scripts/bench_models.py --steps 2 > $OUT/ncu_dp_$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/prof_dp_$TAG.ncu-rep --page raw --csv > $OUT/prof_dp_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_dp_$TAG.ncu-rep --page details > $OUT/prof_dp_${TAG}_details.txt 2>/dev/null
grep -E "Duration|DRAM Throughput|Registers Per|Achieved Occ|Issue Slots Busy|Executed Ipc Active|Avg. Active Threads|Avg. Not Predicated" $OUT/prof_dp_${TAG}_details.txt
