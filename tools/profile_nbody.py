#!/usr/bin/env python
"""Short driver for ncu: fused symplectic N-body kernel (64 bodies x 8192 systems, BABs9o7H, 10 steps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402

systems = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
state, m = de.benchmarks.nbody_ensemble(systems, bodies=64, seed=0)
y0 = torch.as_tensor(np.ascontiguousarray(np.moveaxis(state, 1, -1))).cuda()
a = de.OdeSystem(de.rhs_plugins.nbody, y0, t=(0.0, 0.1), dt=0.01, constants=dict(m=torch.as_tensor(m), eps2=1e-4, G=1.0), ensemble=True)
a.method = "BABS9O7H"
for rep in range(2):
    a.reset()
    a.integrate()
    torch.cuda.synchronize()
print("done", int(a.accepted.sum()))
