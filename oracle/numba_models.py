"""numba restatement of VonMises3D.evaluate -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

SURVEY.md 8(d) asks for compiled CPU baselines beside the reference's own CPython loop so that the GPU
numbers are not compared with an interpreter only: the C port (oracle/fcx_oracle.c, OpenMP) is one, this
`@njit(parallel=True)` loop -- what a maintainer of the reference would most plausibly write first -- is the
other.  Same operation order as the reference (src/fenics_constitutive/models/
mises_plasticity_isotropic_hardening.py:57-175, line numbers in the comments), checked against the C oracle
and the golden fixtures by tests/test_oracle.py.  Nothing in the product imports this module.
"""
from __future__ import annotations

import numpy as np

try:
    from numba import njit, prange
except Exception:  # numba absent: the baseline is simply not available
    njit = None


def available() -> bool:
    return njit is not None


if njit is not None:

    @njit(parallel=True, cache=False, fastmath=False)
    def _mises(params, grad, stress, tangent, eps_n, alpha, flag):
        ka, mu, y0, y00, w = params[0], params[1], params[2], params[3], params[4]
        c23 = np.sqrt(2.0 / 3.0)
        r2 = 1 / 2**0.5  # models/utils.py:199-204
        n = alpha.shape[0]
        failed = 0
        for q in prange(n):
            g = grad[q * 9:q * 9 + 9]
            e = np.empty(6)
            e[0], e[1], e[2] = g[0], g[4], g[8]
            e[3], e[4], e[5] = r2 * (g[1] + g[3]), r2 * (g[2] + g[6]), r2 * (g[5] + g[7])
            sig = stress[q * 6:q * 6 + 6]
            ep = eps_n[q * 6:q * 6 + 6]
            tr_eps = (e[0] + e[1]) + e[2]                                   # :75
            tr_sig = (sig[0] + sig[1]) + sig[2]                             # :81
            ds = np.empty(6)
            st = np.empty(6)
            dot = 0.0
            for k in range(6):
                i2 = 1.0 if k < 3 else 0.0
                ed = e[k] - tr_eps * i2 / 3                                 # :76
                ds[k] = 2 * mu * ed                                         # :79
                st[k] = (sig[k] - tr_sig * i2 / 3) + ds[k]                  # :80-85
                dot += st[k] * st[k]
            nrm = np.sqrt(dot)                                              # :88
            a_n = alpha[q]
            phi = nrm - c23 * (y0 + (y00 - y0) * (1 - np.exp(-w * a_n)))    # :91-94
            xn = np.zeros(6)
            g1 = 0.0
            xc1 = 0.0
            xc2 = 0.0
            if phi > 0:                                                     # :98
                g0 = 1.0
                xr = 1.0
                it = 0
                for k in range(6):
                    xn[k] = st[k] / nrm                                     # :108
                while abs(xr) > 1e-12 and abs(g1 - g0) > 1e-8 * abs(g1):    # :129-131
                    g0 = g1
                    it += 1
                    xr = nrm - 2 * mu * g0 - c23 * (y0 + (y00 - y0) * (1 - np.exp(-w * (a_n + c23 * g0))))
                    xg = -2 * mu - (2.0 / 3.0) * (y00 - y0) * w * np.exp(-w * (a_n + c23 * g0))
                    g1 = g0 - xr / xg                                       # :139
                    if it > 100:                                            # :141-143
                        failed += 1
                        break
                xg = -2 * mu - (2.0 / 3.0) * (y00 - y0) * w * np.exp(-w * (a_n + c23 * g1))  # :147
                xc1 = -1 / xg                                               # :150
                xc2 = g1 / nrm                                              # :151
                flag[q] = 1
            else:
                flag[q] = 0
            for k in range(6):
                ep[k] += g1 * xn[k]                                         # :161
            alpha[q] = a_n + c23 * g1                                       # :162
            for k in range(6):
                i2 = 1.0 if k < 3 else 0.0
                sig[k] += ka * tr_eps * i2 + ds[k] - 2 * mu * g1 * xn[k]    # :165-167
            cpp = 2 * mu * (1 - 2 * mu * xc2)
            cnn = 4 * mu * mu * (xc2 - xc1)
            C = tangent[q * 36:q * 36 + 36]
            for i in range(6):
                for j in range(6):
                    xioi = 1.0 if (i < 3 and j < 3) else 0.0
                    xpp = (1.0 if i == j else 0.0) - (1.0 / 3.0) * xioi
                    C[i * 6 + j] = ka * xioi + cpp * xpp + cnn * (xn[i] * xn[j])  # :170-175
        return failed


class VonMises3D:
    """Same call shape as the reference class; evaluate() runs the numba loop on all numba threads."""

    def __init__(self, param: dict):
        if njit is None:
            raise RuntimeError("numba is not installed")
        self.params = np.array([param["p_ka"], param["p_mu"], param["p_y0"], param["p_y00"], param["p_w"]], dtype=np.float64)
        self.plastic_flag = None

    def evaluate(self, t, del_t, grad_del_u, stress, tangent, history) -> None:
        n = history["alpha"].size
        assert grad_del_u.size == 9 * n and stress.size == 6 * n and tangent.size == 36 * n
        flag = np.zeros(n, dtype=np.uint8)
        failed = _mises(self.params, grad_del_u, stress, tangent, history["eps_n"], history["alpha"], flag)
        self.plastic_flag = flag
        if failed:
            raise RuntimeError("Newton-Raphson method did not converge for plastic multiplier.")
