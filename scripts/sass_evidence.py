#!/usr/bin/env python
"""SASS evidence for profiles/: per-kernel instruction counts from `cuobjdump -sass` of the in-tree objects
(sm_100a cubins of libfcx.so) -- the async-copy machinery (UBLKCP = 1-D bulk async copy / TMA, LDGSTS =
cp.async, SYNCS = mbarrier), the fp64 pipe (DFMA / DMUL / DADD, MUFU for exp / rsqrt seeds) and the plain
memory instructions -- plus the first lines that use each of the async mnemonics.

    python scripts/sass_evidence.py            # writes profiles/sass_<kernel>.txt and profiles/sass_counts.json
"""
from __future__ import annotations

import collections
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "fenics_constitutive_b200", "csrc", "obj")
OUT = os.path.join(ROOT, "profiles")
MNEMONICS = ["UBLKCP", "LDGSTS", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU", "LDG", "STG", "LDS", "STS", "LDC", "ULDC",
             "BAR", "ATOM", "RED", "HMMA", "UTCMMA", "QMMA"]
KERNELS = {  # file tag -> (object, regex on the demangled name)
    "mises_ostage_64_8_tangent": ("fcx_api.o", r"fcx_mises_ostage_kernel<64, 8, true>"),
    "mises_tile_128_stress_only": ("fcx_api.o", r"fcx_tile_kernel<fcx::MisesModel<false>, 128, false>"),
    "elastic_full_tile_128": ("fcx_api.o", r"fcx_tile_kernel<fcx::ElasticModel<6, 3>, 128, true>"),
    "kelvin_full_tile_128": ("fcx_api.o", r"fcx_tile_kernel<fcx::KelvinModel<6, 3>, 128, true>"),
    "maxwell_full_tile_128": ("fcx_api.o", r"fcx_tile_kernel<fcx::MaxwellModel<6, 3>, 128, true>"),
    "drucker_prager_classic_tile_128": ("fcx_api.o", r"fcx_tile_kernel<fcx::DruckerPragerModel<false, 1>, 128, true>"),
    "drucker_prager_classic_reference_spelling_tile_128": ("fcx_api.o", r"fcx_tile_kernel<fcx::DruckerPragerModel<false, 0>, 128, true>"),
    "mises_form_p2_q2": ("fcx_api.o", r"fcx_mises_form_kernel<10, 4, 64, 8>"),
    "gather_staged_p2_q2": ("fcx_gather.o", r"gather_staged_kernel<3, 10, 4, false>"),
    "gather_cell_p2": ("fcx_gather.o", r"gather_cell_kernel<10, false>"),
    "tangent_apply_rec_p2_q2": ("fcx_assemble.o", r"qp_cell_kernel<3, 6, 10, 4, 3>"),
    "krylov_gsum_dots": ("fcx_krylov.o", r"gsum_dots_kernel<3>"),
    "krylov_cg_update": ("fcx_krylov.o", r"cg_update_kernel"),
    "krylov_halo_push": ("fcx_krylov.o", r"halo_push_kernel<3>"),
    "map_rows_to_sub": ("fcx_maps.o", r"map_rows_kernel<false, double2>"),
    "wire_pack": ("fcx_host.o", r"wire_pack_kernel"),
}


def functions(obj):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : (\S+)\n", txt)
    names = parts[1::2]
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return list(zip(dem, parts[2::2]))


def main():
    os.makedirs(OUT, exist_ok=True)
    cache, counts_all = {}, {}
    for tag, (obj, pat) in KERNELS.items():
        if obj not in cache:
            cache[obj] = functions(obj)
        hit = [(n, body) for n, body in cache[obj] if re.search(pat, n)]
        if not hit:
            print("not found:", tag, pat)
            continue
        name, body = hit[0]
        lines = [ln for ln in body.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
        ops = collections.Counter()
        first = {}
        for ln in lines:
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if not m:
                continue
            base = m.group(1).split(".")[0]
            for mn in MNEMONICS:
                if base == mn or (mn in ("ATOM", "RED", "BAR") and base.startswith(mn)):
                    ops[mn] += 1
                    first.setdefault(mn, re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", ln).strip())
        counts_all[tag] = {"kernel": name.split("(")[0], "instructions": len(lines), **{k: ops[k] for k in MNEMONICS if ops[k]}}
        with open(os.path.join(OUT, f"sass_{tag}.txt"), "w") as f:
            f.write(f"# cuobjdump -sass fenics_constitutive_b200/csrc/obj/{obj}  (sm_100a), kernel:\n# {name}\n")
            f.write(f"# {len(lines)} SASS instructions; counts of the mnemonics that matter here:\n")
            for k in MNEMONICS:
                if ops[k]:
                    f.write(f"#   {k:7s} {ops[k]:5d}   first: {first[k]}\n")
            f.write("# no tensor-core / TMEM instructions (HMMA / UTCMMA / QMMA): nothing on this path is a dense contraction\n"
                    if not (ops["HMMA"] or ops["UTCMMA"] or ops["QMMA"]) else "")
            f.write("\n# async-copy, mbarrier and fp64 lines in program order (first 60):\n")
            shown = 0
            for ln in lines:
                if re.search(r"UBLKCP|LDGSTS|SYNCS|DFMA|MUFU", ln) and shown < 60:
                    f.write(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", ln).rstrip() + "\n")
                    shown += 1
    with open(os.path.join(OUT, "sass_counts.json"), "w") as f:
        json.dump(counts_all, f, indent=1)
    for tag, c in counts_all.items():
        print(tag, {k: v for k, v in c.items() if k != "kernel"})


if __name__ == "__main__":
    main()
