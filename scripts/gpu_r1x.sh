#!/bin/bash
# Round-1 session x: vectorised PCG kernels, constant-tangent host pipeline: parity, Newton solve, host e2e of elastic / Kelvin.
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_r1x.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r1x.log
timeout 900 python scripts/bench_newton.py --n 55 --steps 2 > $OUT/newton55_r1x.log 2>&1; echo "newton rc=$?"; tail -1 $OUT/newton55_r1x.log
python - > $OUT/e2e_models_r1x.log 2>&1 <<'PY'
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
        for wire in (0, 1):
            L.fcx_host_wire(wire)
            r1 = best(lambda: el.evaluate(0.0, 1.0, a[0].numpy(), a[1].numpy(), a[2].numpy(), None))
            r2 = best(lambda: kv.evaluate(0.0, 2.0, a[0].numpy(), a[1].numpy(), a[2].numpy(), {"strain_visco": a[3].numpy(), "strain": a[4].numpy()}))
            print(json.dumps({"constraint": cons.name, "pinned": pinned, "wire": wire, "elastic_MQPs": r1, "kelvin_MQPs": r2}), flush=True)
        del a
PY
echo "e2e models rc=$?"; cat $OUT/e2e_models_r1x.log
