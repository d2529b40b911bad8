"""Generate tests/golden/adapters.npz by running the reference's UNMODIFIED 3D -> 1D/2D adapters
(src/fenics_constitutive/models/utils.py:211-412) around its own FULL models.

Run in the build container only (needs /root/reference):
    python -m oracle.gen_golden_adapters
Three consecutive calls per case on the SAME adapter object and the same caller arrays, so the
fixtures also pin the reference's persistent 3D scratch arrays (the unmapped
3D stress components accumulate across calls).  Cases: UniaxialStrainFrom3D / PlaneStrainFrom3D around
LinearElasticityModel(FULL) and VonMises3D (history passes through as the 3D model's own arrays).
"""
from __future__ import annotations

import os

import numpy as np

from . import ref_shim
from .gen_golden import ELASTIC_PARAMS, MISES_PARAMS, OUT


def main() -> None:
    m = ref_shim.load()
    C = m.StressStrainConstraint
    rng = np.random.default_rng(20261018)
    n, ncalls = 64, 3
    out = {"n": np.array(n), "ncalls": np.array(ncalls)}
    for aname, adapter, g, s, scale in (("uniaxial_strain", m.UniaxialStrainFrom3D, 1, 1, 8e-3),
                                        ("plane_strain", m.PlaneStrainFrom3D, 2, 4, 3e-3)):
        for lname in ("elastic", "mises"):
            law3d = m.LinearElasticityModel(ELASTIC_PARAMS, C.FULL) if lname == "elastic" else m.VonMises3D(MISES_PARAMS)
            wrap = adapter(law3d)
            assert wrap.constraint.geometric_dim == g and wrap.constraint.stress_strain_dim == s
            f = (1e-3 / scale) if lname == "elastic" else 1.0
            stress = np.zeros(n * s)
            tangent = np.full(n * s * s, np.nan)
            history = None
            if lname == "mises":
                history = {"eps_n": np.zeros(n * 6), "alpha": np.zeros(n)}
            key = f"{aname}_{lname}"
            for k in range(ncalls):
                grad = rng.standard_normal(n * g * g) * scale * f
                wrap.evaluate(0.0, 1.0, grad, stress, tangent, history)
                out[f"{key}_grad{k}"] = grad
                out[f"{key}_stress{k}"] = stress.copy()
                out[f"{key}_tangent{k}"] = tangent.copy()
                if history is not None:
                    out[f"{key}_eps_n{k}"] = history["eps_n"].copy()
                    out[f"{key}_alpha{k}"] = history["alpha"].copy()
            if history is not None:
                frac = float((history["alpha"] > 0).mean())
                assert 0.1 < frac < 0.95, (key, frac)
                out[f"{key}_plastic_fraction"] = np.array(frac)
    os.makedirs(OUT, exist_ok=True)
    np.savez(os.path.join(OUT, "adapters.npz"), **out)
    print("wrote", os.path.join(OUT, "adapters.npz"), {k: float(v) for k, v in out.items() if k.endswith("fraction")})


if __name__ == "__main__":
    main()
