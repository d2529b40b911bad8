#!/usr/bin/env python
"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenics_constitutive_b200 import synthetic  # noqa: E402
from fenics_constitutive_b200.models import (  # noqa: E402
    LinearElasticityModel, SpringKelvinModel, SpringMaxwellModel, StressStrainConstraint, VonMises3D)

C = StressStrainConstraint
dev = torch.device("cuda", 0)
for n in (1, 128, 1000, 5 * 128 * 148 + 77):
    g, s0, e0, a0 = synthetic.mises_inputs_numpy(n, seed=3)
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    law = VonMises3D(synthetic.MISES_PARAMS)
    law.record_plastic_flag = True
    tg = torch.empty(n * 36, dtype=torch.float64, device=dev)
    law.evaluate(0.0, 1.0, t(g), t(s0), tg, {"eps_n": t(e0), "alpha": t(a0)})
    for cons in (C.UNIAXIAL_STRESS, C.PLANE_STRAIN, C.FULL):
        gd, sd = cons.geometric_dim, cons.stress_strain_dim
        gr = torch.randn(n * gd * gd, dtype=torch.float64, device=dev) * 1e-3
        z = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731
        tg = torch.empty(n * sd * sd, dtype=torch.float64, device=dev)
        LinearElasticityModel(synthetic.ELASTIC_PARAMS, cons).evaluate(0.0, 1.0, gr, z(n * sd), tg, None)
        for cls in (SpringKelvinModel, SpringMaxwellModel):
            cls(synthetic.VISCO_PARAMS, cons).evaluate(
                0.0, 2.0, gr, z(n * sd), tg, {"strain_visco": z(n * sd), "strain": z(n * sd)})
torch.cuda.synchronize()
print("sanitize_small: done")
