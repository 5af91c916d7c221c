#!/usr/bin/env python
"""Short driver for ncu: the fused kernel's Van der Pol instantiation (BASELINE config 4 at reduced size, 2^21 trajectories)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
cfg = de.benchmarks.VDP
y0, mu = de.benchmarks.vdp_ensemble(n)
a = de.OdeSystem(de.rhs_plugins.van_der_pol, torch.as_tensor(y0).cuda(), t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"],
                 atol=cfg["atol"], constants=dict(mu=torch.as_tensor(mu).cuda()), ensemble=True)
a.method = "RK45"
for rep in range(2):
    a.reset()
    a.integrate()
    torch.cuda.synchronize()
print("done", a.stats)
