#!/usr/bin/env python
"""Host-array path under compute-sanitizer: staged pipeline + download wire (slot records, direct
tangent stores into page-locked arrays, plain D2H) on small batches, ragged chunks, several slots;
VonMises3D and the comfe-rs plastic mirrors."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenics_constitutive_b200 import models as M  # noqa: E402
from fenics_constitutive_b200 import synthetic  # noqa: E402
from fenics_constitutive_b200._lib import lib  # noqa: E402

L = lib()
L.fcx_host_chunk_qps(5000)
rs = {"mu": np.array([80769.0]), "kappa": np.array([175000.0])}
for slots in (2, 6):
    L.fcx_host_slots(slots)
    for wire in (2, 1, 0):
        L.fcx_host_wire(wire)
        for pinned in (False, True):
            n = 23_456 + 17 * slots
            g, s0, e0, a0 = synthetic.mises_inputs_numpy(n, seed=5)
            arrs = [g, s0, np.zeros(n * 36), e0, a0, np.zeros(n * 7)]
            keep = []
            if pinned:
                keep = [torch.from_numpy(a).pin_memory() for a in arrs]
                arrs = [t.numpy() for t in keep]
            law = M.VonMises3D(synthetic.MISES_PARAMS)
            law.record_plastic_flag = True
            law.evaluate(0.0, 1.0, arrs[0], arrs[1], arrs[2], {"eps_n": arrs[3], "alpha": arrs[4]})
            lin = M.MisesPlasticityLinearHardening3D({**rs, "y_0": np.array([1200.0]), "h": np.array([200.0])})
            lin.evaluate(0.0, 1.0, arrs[0], arrs[1], arrs[2], {"history": arrs[5]})
            dp = M.DruckerPragerHyperbolic3D({**rs, "a": np.array([300.0]), "b": np.array([0.05]),
                                              "b_flow": np.array([0.01]), "d": np.array([40.0])})
            arrs[1][:] = 0.0
            arrs[5][:] = 0.0
            gd = arrs[0].reshape(n, 9) * 0.5
            gd[:, [0, 4, 8]] *= 0.2
            dp.evaluate(0.0, 1.0, np.ascontiguousarray(gd.ravel()), arrs[1], arrs[2], {"history": arrs[5]})
L.fcx_host_slots(6)
L.fcx_host_wire(1)
print("sanitize_host: done")
