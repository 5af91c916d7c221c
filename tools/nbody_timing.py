#!/usr/bin/env python
"""Timing of BASELINE config 3: gravitational N-body, 64 bodies x 65536 systems, BABs9o7H, dt=0.01, 10 steps."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402

systems = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
state, m = de.benchmarks.nbody_ensemble(systems, bodies=64, seed=0)
y0 = torch.as_tensor(np.ascontiguousarray(np.moveaxis(state, 1, -1))).cuda()
for method in ("BABS9O7H", "ABAS5O6H"):
    a = de.OdeSystem(de.rhs_plugins.nbody, y0, t=(0.0, 0.1), dt=0.01, constants=dict(m=torch.as_tensor(m), eps2=1e-4, G=1.0),
                     ensemble=True)
    a.method = method
    for rep in range(3):
        a.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a.integrate()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    steps = int(a.accepted.sum())
    kicks = int((a.integrator.tableau_intermediate[:, 2] != 0).sum())
    pair_evals = steps * kicks * 64 * 64
    print("%s: %d systems, %.2f ms, %.3e system-steps/s, %.3e pair interactions/s, steps/system %d" %
          (method, systems, ms, steps / ms * 1e3, pair_evals / ms * 1e3, steps // systems))
    yf = a[-1].y
    print("   checksum %016x" % (int(yf.view(torch.int64).to(torch.float64).sum().item()) & 0xffffffffffffffff),
          "sum|y| %.17g" % float(yf.abs().sum()))
