#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export into a small JSON for profiles/.
    python scripts/ncu_extract.py raw.csv out.json [--traffic profiles/mises_traffic.json --qps N]"""
import argparse
import csv
import json

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "sm__cycles_elapsed.max",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}

ap = argparse.ArgumentParser()
ap.add_argument("raw")
ap.add_argument("out")
ap.add_argument("--traffic", default=None)
ap.add_argument("--qps", type=int, default=16_000_000)
ap.add_argument("--bytes-per-qp", type=int, default=568)
ap.add_argument("--source", default="")
a = ap.parse_args()
rows = list(csv.reader(open(a.raw)))
hdr, units = rows[0], rows[1]
out = []
for vals in rows[2:]:
    m = {}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            m[w] = {"unit": units[i], "value": vals[i]}
    out.append({"kernel": vals[hdr.index("Kernel Name")], "metrics": m})
json.dump(out if len(out) != 1 else out[0], open(a.out, "w"), indent=1)
if a.traffic and out:
    m = out[0]["metrics"]
    rd = float(m["dram__bytes_read.sum"]["value"]) * UNIT[m["dram__bytes_read.sum"]["unit"]]
    wr = float(m["dram__bytes_write.sum"]["value"]) * UNIT[m["dram__bytes_write.sum"]["unit"]]
    json.dump({"kernel": out[0]["kernel"], "qps": a.qps, "dram_bytes_read": rd, "dram_bytes_write": wr,
               "dram_bytes_per_launch": rd + wr, "algorithmic_bytes_per_launch": a.bytes_per_qp * a.qps,
               "source": a.source}, open(a.traffic, "w"), indent=1)
