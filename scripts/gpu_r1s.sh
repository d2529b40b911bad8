#!/bin/bash
# Round-1 session s: packed wire with streaming stores: parity of the host path, e2e with wire on/off, threads sweep.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "memory_kinds or golden or sizes" > $OUT/pytest_r1s.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_r1s.log
python - > $OUT/e2e_sweep_r1s.log 2>&1 <<'PY'
import json, time, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fenics_constitutive_b200 import synthetic
from fenics_constitutive_b200._lib import lib
from fenics_constitutive_b200.models import VonMises3D
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
for wire in (0, 1):
    L.fcx_host_wire(wire)
    for thr in (4, 8, 12, 14, 16):
        L.fcx_host_threads(thr)
        for chunk in (1 << 15, 1 << 16, 1 << 17):
            L.fcx_host_chunk_qps(chunk)
            print(json.dumps({"wire": wire, "threads": thr, "chunk": chunk, "pinned_MQPs": round(run(h), 1), "pageable_MQPs": round(run(pg), 1)}), flush=True)
        if wire == 0 and thr >= 8: break
PY
echo "sweep rc=$?"; cat $OUT/e2e_sweep_r1s.log
