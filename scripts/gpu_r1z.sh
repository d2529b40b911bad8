#!/bin/bash
# Round-1 session z: direct download wire (plastic tangents stored in place in page-locked caller arrays),
# generic plastic wire for the comfe-rs mirrors: parity, e2e sweep, bench.
OUT=gpurun_out; mkdir -p $OUT
nproc > $OUT/gpu_r1z.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $OUT/gpu_r1z.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_r1z.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r1z.log
python - > $OUT/e2e_sweep_r1z.log 2>&1 <<'PY'
import json, time, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fenics_constitutive_b200 import synthetic
from fenics_constitutive_b200._lib import lib
from fenics_constitutive_b200.models import VonMises3D, MisesPlasticityLinearHardening3D
L = lib()
n = 16_000_000
pin = lambda m: torch.empty(m, dtype=torch.float64).pin_memory()
h = [pin(n*9), pin(n*6), pin(n*6), pin(n), pin(n*36)]
pg = [torch.from_numpy(np.zeros(m)) for m in (n*9, n*6, n*6, n, n*36)]
rng = np.random.default_rng(99)
gr = rng.standard_normal(n*9) * synthetic.MISES_GRAD_STD
h[0].numpy()[:] = gr; pg[0].numpy()[:] = gr
law = VonMises3D(synthetic.MISES_PARAMS)
def run(arrs, reps=3):
    best = 0
    for i in range(reps + 1):
        for a in arrs[1:4]: a.zero_()
        t0 = time.perf_counter()
        law.evaluate(0.0, 1.0, arrs[0].numpy(), arrs[1].numpy(), arrs[4].numpy(), {"eps_n": arrs[2].numpy(), "alpha": arrs[3].numpy()})
        dt = time.perf_counter() - t0
        if i > 0: best = max(best, n / dt)
    return best / 1e6
for wire, thr, chunk in ((0, 14, 1 << 17), (1, 14, 1 << 16), (2, 14, 1 << 16), (2, 8, 1 << 16), (2, 4, 1 << 16), (2, 16, 1 << 16),
                         (2, 14, 1 << 15), (2, 14, 1 << 17), (2, 14, 1 << 18)):
    L.fcx_host_wire(wire); L.fcx_host_threads(thr); L.fcx_host_chunk_qps(chunk)
    row = {"wire": wire, "threads": thr, "chunk": chunk, "pinned_MQPs": round(run(h), 1)}
    if wire < 2 or (thr == 14 and chunk == 1 << 16):
        row["pageable_MQPs"] = round(run(pg), 1)
    print(json.dumps(row), flush=True)
L.fcx_host_threads(14); L.fcx_host_chunk_qps(1 << 18)
# comfe-rs linear-hardening Mises through the same wire (one [n][7] history array)
lin = MisesPlasticityLinearHardening3D({"mu": np.array([80769.0]), "kappa": np.array([175000.0]), "y_0": np.array([1200.0]), "h": np.array([200.0])})
hh = pin(n * 7); ph = torch.from_numpy(np.zeros(n * 7))
def run_lin(arrs, hist, reps=2):
    best = 0
    for i in range(reps + 1):
        arrs[1].zero_(); hist.zero_()
        t0 = time.perf_counter()
        lin.evaluate(0.0, 1.0, arrs[0].numpy(), arrs[1].numpy(), arrs[4].numpy(), {"history": hist.numpy()})
        dt = time.perf_counter() - t0
        if i > 0: best = max(best, n / dt)
    return best / 1e6
for wire in (0, 1, 2):
    L.fcx_host_wire(wire)
    print(json.dumps({"model": "MisesPlasticityLinearHardening3D", "wire": wire, "pinned_MQPs": round(run_lin(h, hh), 1), "pageable_MQPs": round(run_lin(pg, ph), 1)}), flush=True)
PY
echo "sweep rc=$?"; cat $OUT/e2e_sweep_r1z.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_r1z.json 2> $OUT/bench_r1z.err; echo "bench rc=$?"; cat $OUT/bench_r1z.json; tail -3 $OUT/bench_r1z.err
