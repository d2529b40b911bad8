#!/bin/bash
# Round-1 session zi: staged gather, second version (table in constant memory, lane = cell): parity, sanitizer, A/B, ncu.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gather.py tests/test_solver_gpu.py -m gpu -x -q > $OUT/pytest_r1zi.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r1zi.log
timeout 600 python scripts/bench_gather.py --ctas 0,4,5,6,8 > $OUT/gather_ab_r1zi.jsonl 2> $OUT/gather_ab_r1zi.err; echo "gather rc=$?"; grep '"variant": 1' $OUT/gather_ab_r1zi.jsonl; grep '"variant": 0, "ctas_per_sm": 0' $OUT/gather_ab_r1zi.jsonl; tail -3 $OUT/gather_ab_r1zi.err
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > $OUT/racecheck_r1zi.log 2>&1; echo "racecheck rc=$?"; tail -2 $OUT/racecheck_r1zi.log
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > $OUT/memcheck_r1zi.log 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/memcheck_r1zi.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:gather_staged -c 2 -o $OUT/prof_gather_r1zi python scripts/bench_gather.py --reps 2 --ctas 0 > $OUT/ncu_gather_r1zi.log 2>&1; echo "ncu gather rc=$?"
ncu -i $OUT/prof_gather_r1zi.ncu-rep --page raw --csv > $OUT/prof_gather_r1zi_raw.csv 2>/dev/null
python scripts/ncu_extract.py $OUT/prof_gather_r1zi_raw.csv $OUT/r1zi_gather_ncu_full.json
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/prof_gather_r1zi_raw.csv')))
hdr,v=rows[0],rows[2]
for k in ['gpu__time_duration.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum','smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum']:
    if k in hdr: print(k, v[hdr.index(k)])
PY
