#!/usr/bin/env python
"""Short driver for ncu: the DENSE instantiation of the fused kernel (one node per accepted step into the paged node store),
2^18 Lorenz trajectories, followed by the index build and one 8 x N grid evaluation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
cfg = de.benchmarks.LORENZ
y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(n)).cuda()
a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"], ensemble=True,
                 dense_output=True)
a.method = "RK45"
for rep in range(2):
    a.reset()
    a.integrate()
    torch.cuda.synchronize()
g = a.sol.grid(torch.linspace(0.1, 1.9, 8, dtype=torch.float64, device="cuda"))
torch.cuda.synchronize()
print("done", a.stats, tuple(g.shape))
