#!/bin/bash
# Round-1 session zk: inexact Newton (Eisenstat-Walker forcing) in the stand-in solver; host e2e of the constant-tangent
# models with the new pipeline defaults (6 slots, asynchronous expansion, pool size by kind of work).
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_r1zk.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r1zk.log
timeout 900 python scripts/bench_newton.py --n 55 --steps 2 --forcing ew > $OUT/newton55_ew_r1zk.log 2>&1; echo "newton ew rc=$?"; tail -1 $OUT/newton55_ew_r1zk.log
python - > $OUT/e2e_models_r1zk.log 2>&1 <<'PY'
import json, time, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fenics_constitutive_b200 import synthetic
from fenics_constitutive_b200._lib import lib
from fenics_constitutive_b200.models import LinearElasticityModel, SpringKelvinModel, StressStrainConstraint as C
L = lib()
n = 16_000_000
def arrays(sizes, pinned):
    return [torch.zeros(m, dtype=torch.float64).pin_memory() if pinned else torch.from_numpy(np.zeros(m)) for m in sizes]
def best(fn, reps=3):
    b = 0
    for i in range(reps + 1):
        t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
        if i > 0: b = max(b, n / dt)
    return round(b / 1e6, 1)
for cons in (C.FULL, C.PLANE_STRAIN):
    g, s = cons.geometric_dim, cons.stress_strain_dim
    for pinned in (True, False):
        a = arrays([n*g*g, n*s, n*s*s, n*s, n*s], pinned)
        a[0].numpy()[:] = np.random.default_rng(1).standard_normal(n*g*g) * 1e-4
        el = LinearElasticityModel(synthetic.ELASTIC_PARAMS, cons)
        kv = SpringKelvinModel(synthetic.VISCO_PARAMS, cons)
        for thr in (0, 8, 14):
            if thr: L.fcx_host_threads(thr)
            r1 = best(lambda: el.evaluate(0.0, 1.0, a[0].numpy(), a[1].numpy(), a[2].numpy(), None))
            r2 = best(lambda: kv.evaluate(0.0, 2.0, a[0].numpy(), a[1].numpy(), a[2].numpy(), {"strain_visco": a[3].numpy(), "strain": a[4].numpy()}))
            print(json.dumps({"constraint": cons.name, "pinned": pinned, "threads": thr or "default", "elastic_MQPs": r1, "kelvin_MQPs": r2}), flush=True)
        del a
PY
echo "e2e models rc=$?"; cat $OUT/e2e_models_r1zk.log
