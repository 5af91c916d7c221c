#!/usr/bin/env python
"""Resident blocks per SM of the fused kernel against ensemble size: time of one Lorenz RK45CK pass for every
(n, DSB_ERK_BLOCKS_PER_SM) pair, and with the library's own choice (grid_for_ensemble, csrc/plan.h)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402

cfg = de.benchmarks.LORENZ
out = {}
for logn in (13, 14, 15, 16, 17, 18, 19, 20):
    n = 1 << logn
    y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(1 << 20)[:, ::(1 << 20) // n].copy()).cuda()
    a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"], ensemble=True)
    a.method = "RK45"
    a.lazy_status = True
    row = {}
    for w in ("auto", 1, 2, 3, 4, 5, 6):
        if w == "auto":
            os.environ.pop("DSB_ERK_BLOCKS_PER_SM", None)
        else:
            os.environ["DSB_ERK_BLOCKS_PER_SM"] = str(w)
        best = 1e30
        for rep in range(5):
            a.reset()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a.integrate()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        row[str(w)] = round(best, 4)
    out[n] = row
    print(n, row, flush=True)
os.environ.pop("DSB_ERK_BLOCKS_PER_SM", None)
print(json.dumps(out))
