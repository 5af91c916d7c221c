#!/usr/bin/env python
"""Short driver for ncu: the fused kernel instantiated for a caller-defined right-hand side (the Roessler plug-in of
desolver_b200.benchmarks, built by rhs_plugins.register_device_rhs), 2^20 trajectories, t in [0, 10]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
rng = np.random.default_rng(3)
y0 = torch.as_tensor(rng.uniform([-8, -8, 0], [8, 8, 12], (n, 3)).T.copy()).cuda()
a = de.OdeSystem(de.benchmarks.rossler_plugin(), y0, t=(0.0, 10.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
a.method = "RK45"
for rep in range(2):
    a.reset()
    a.integrate()
    torch.cuda.synchronize()
print("done", a.stats)
