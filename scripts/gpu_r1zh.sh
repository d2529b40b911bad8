#!/bin/bash
# Round-1 session zh: ncu of the staged gather kernel (what bounds it now?)
OUT=gpurun_out; mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:gather_staged -c 2 -o $OUT/prof_gather_r1zh python scripts/bench_gather.py --reps 2 --ctas 0 > $OUT/ncu_gather_r1zh.log 2>&1; echo "ncu gather rc=$?"; tail -3 $OUT/ncu_gather_r1zh.log
ncu -i $OUT/prof_gather_r1zh.ncu-rep --page raw --csv > $OUT/prof_gather_r1zh_raw.csv 2>/dev/null
ncu -i $OUT/prof_gather_r1zh.ncu-rep --page details > $OUT/prof_gather_r1zh_details.txt 2>/dev/null
python scripts/ncu_extract.py $OUT/prof_gather_r1zh_raw.csv $OUT/r1zh_gather_ncu_full.json; grep -E "Pipe|pipe|Stall|stall|L1/TEX Hit|Bank|bank|Issue Slots|Mem Busy|Max Bandwidth|Mem Pipes" $OUT/prof_gather_r1zh_details.txt | head -60
