"""Generate tests/golden/*.npz by running the reference's UNMODIFIED models.

Run in the build container only (needs /root/reference):
    python -m oracle.gen_golden
The fixtures hold seeded inputs and the reference's outputs for every model
on the north-star list; tests compare the C oracle (CPU) and the CUDA path
(GPU) against them.  Parameters follow the reference's own tests:
  elasticity  E=42, nu=0.3           tests/models/test_elasticity.py:22-23
  Mises       ka=175000 mu=80769 y0=1200 y00=2500 w=200
                                     tests/models/test_plasticity.py:19-25
  visco       E0=42 E1=10 tau=10 nu=0.2, dt=2
                                     tests/models/test_viscoelasticity.py:20-23,50
"""
from __future__ import annotations

import os

import numpy as np

from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

ELASTIC_PARAMS = {"E": 42.0, "nu": 0.3}
MISES_PARAMS = {"p_ka": 175000.0, "p_mu": 80769.0, "p_y0": 1200.0, "p_y00": 2500.0, "p_w": 200.0}
VISCO_PARAMS = {"E0": 42.0, "E1": 10.0, "tau": 10.0, "nu": 0.2}
CONSTRAINTS = ["UNIAXIAL_STRAIN", "UNIAXIAL_STRESS", "PLANE_STRAIN", "PLANE_STRESS", "FULL"]


def main() -> None:
    m = ref_shim.load()
    C = m.StressStrainConstraint
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261017)

    # --- Mandel conversion known answers + random -------------------------
    conv = {}
    for name in CONSTRAINTS:
        c = C[name]
        g = c.geometric_dim
        grad = rng.standard_normal(37 * g * g)
        conv[f"{name}_grad"] = grad
        conv[f"{name}_strain"] = m.strain_from_grad_u(grad, c)
        conv[f"{name}_D"] = m.get_elastic_tangent(42.0, 0.3, c)
        conv[f"{name}_I2"] = m.get_identity(c.stress_strain_dim, c)
    conv["lame_42_0.3"] = np.array(m.lame_parameters(42.0, 0.3))
    np.savez(os.path.join(OUT, "conversions.npz"), **conv)

    # --- LinearElasticityModel, all constraints ---------------------------
    el = {}
    n = 96
    for name in CONSTRAINTS:
        c = C[name]
        g, s = c.geometric_dim, c.stress_strain_dim
        law = m.LinearElasticityModel(ELASTIC_PARAMS, c)
        grad = rng.standard_normal(n * g * g) * 1e-3
        stress0 = rng.standard_normal(n * s) * 0.1
        stress = stress0.copy()
        tangent = np.full(n * s * s, np.nan)
        law.evaluate(0.0, 1.0, grad, stress, tangent, None)
        el[f"{name}_grad"] = grad
        el[f"{name}_stress_in"] = stress0
        el[f"{name}_stress_out"] = stress
        el[f"{name}_tangent"] = tangent
    np.savez(os.path.join(OUT, "elasticity.npz"), **el)

    # --- VonMises3D: two increments, ~50 % plastic ------------------------
    n = 384
    law = m.VonMises3D(MISES_PARAMS)
    mi = {}
    stress = np.zeros(n * 6)
    eps_n = np.zeros(n * 6)
    alpha = np.zeros(n)
    for step in range(2):
        grad = rng.standard_normal(n * 9) * 2.906e-3
        tangent = np.full(n * 36, np.nan)
        mi[f"s{step}_grad"] = grad
        mi[f"s{step}_stress_in"] = stress.copy()
        mi[f"s{step}_eps_n_in"] = eps_n.copy()
        mi[f"s{step}_alpha_in"] = alpha.copy()
        a0 = alpha.copy()
        law.evaluate(0.0, 1.0, grad, stress, tangent, {"eps_n": eps_n, "alpha": alpha})
        mi[f"s{step}_stress_out"] = stress.copy()
        mi[f"s{step}_eps_n_out"] = eps_n.copy()
        mi[f"s{step}_alpha_out"] = alpha.copy()
        mi[f"s{step}_tangent"] = tangent
        mi[f"s{step}_plastic"] = (alpha > a0).astype(np.uint8)
    np.savez(os.path.join(OUT, "mises.npz"), **mi)

    # --- Kelvin / Maxwell, all constraints, three increments --------------
    for cls_name in ("SpringKelvinModel", "SpringMaxwellModel"):
        vi = {}
        n = 64
        for name in CONSTRAINTS:
            c = C[name]
            g, s = c.geometric_dim, c.stress_strain_dim
            law = getattr(m, cls_name)(VISCO_PARAMS, c)
            stress = rng.standard_normal(n * s) * 0.05
            ev = rng.standard_normal(n * s) * 1e-4
            et = rng.standard_normal(n * s) * 1e-3
            vi[f"{name}_stress_in"] = stress.copy()
            vi[f"{name}_strain_visco_in"] = ev.copy()
            vi[f"{name}_strain_in"] = et.copy()
            for step, dt in enumerate((1e-8, 2.0, 0.1)):
                grad = rng.standard_normal(n * g * g) * 1e-3
                tangent = np.full(n * s * s, np.nan)
                law.evaluate(0.0, dt, grad, stress, tangent, {"strain_visco": ev, "strain": et})
                vi[f"{name}_s{step}_dt"] = np.array(dt)
                vi[f"{name}_s{step}_grad"] = grad
                vi[f"{name}_s{step}_stress_out"] = stress.copy()
                vi[f"{name}_s{step}_strain_visco_out"] = ev.copy()
                vi[f"{name}_s{step}_strain_out"] = et.copy()
                vi[f"{name}_s{step}_tangent"] = tangent
        fname = "kelvin.npz" if "Kelvin" in cls_name else "maxwell.npz"
        np.savez(os.path.join(OUT, fname), **vi)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
