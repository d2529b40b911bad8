#!/usr/bin/env python
"""Host-array path under compute-sanitizer: staged pipeline + packed wire on a small batch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenics_constitutive_b200 import synthetic  # noqa: E402
from fenics_constitutive_b200._lib import lib  # noqa: E402
from fenics_constitutive_b200.models import VonMises3D  # noqa: E402

L = lib()
L.fcx_host_chunk_qps(5000)
for wire in (1, 0):
    L.fcx_host_wire(wire)
    n = 23_456
    g, s0, e0, a0 = synthetic.mises_inputs_numpy(n, seed=5)
    law = VonMises3D(synthetic.MISES_PARAMS)
    law.record_plastic_flag = True
    law.evaluate(0.0, 1.0, g, s0, np.zeros(n * 36), {"eps_n": e0, "alpha": a0})
print("sanitize_host: done")
