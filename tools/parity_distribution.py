#!/usr/bin/env python
"""Distribution of the GPU-vs-oracle state error over a large Lorenz sample (fast and strict arithmetic)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402
from desolver_b200.integrators import tableaux as tb  # noqa: E402
from oracle import oracle as orc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
cfg = de.benchmarks.LORENZ
y0 = de.benchmarks.lorenz_ensemble(n)
t0 = time.time()
ref = orc.erk_ensemble(orc.RHS_LORENZ, y0, 0.0, cfg["tf"], cfg["dt0"], 1e-9, 1e-9, *tb.rk_arrays("RK45CKSolver"),
                       params=[10.0, 28.0, 8.0 / 3.0], threads=os.cpu_count())
print("oracle: %d trajectories in %.1f s" % (n, time.time() - t0))
for strict in (False, True):
    a = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(0.0, cfg["tf"]), dt=cfg["dt0"], rtol=1e-9, atol=1e-9, ensemble=True,
                     strict=strict)
    a.integrate()
    y = a[-1].y.cpu().numpy()
    err = np.max(np.abs(y - ref["y"]), axis=0) / np.max(np.abs(ref["y"]), axis=0)
    same = (a.accepted.cpu().numpy() == ref["accepted"]) & (a.rejected.cpu().numpy() == ref["rejected"])
    print("strict=%s: counts identical %.6f (%d differ), rel err max %.3g p99.99 %.3g p99.9 %.3g p99 %.3g median %.3g, > 1e-10: %d of %d" %
          (strict, same.mean(), int((~same).sum()), err.max(), np.percentile(err, 99.99), np.percentile(err, 99.9), np.percentile(err, 99),
           np.median(err), int((err > 1e-10).sum()), n))
    worst = np.argsort(err)[-3:]
    print("   worst:", [(int(i), float(err[i]), int(ref["accepted"][i]), bool(same[i])) for i in worst])
