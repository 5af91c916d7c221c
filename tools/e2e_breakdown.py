#!/usr/bin/env python
"""Where the end-to-end (host buffers) time of one Lorenz 2^20 pass goes: CUDA events between the phases."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402

n = 1 << 20
cfg = de.benchmarks.LORENZ
y0_pinned = torch.from_numpy(de.benchmarks.lorenz_ensemble(n)).pin_memory()
out_y = torch.empty((3, n), dtype=torch.float64).pin_memory()
out_c = torch.empty((3, n), dtype=torch.int32).pin_memory()
dev = torch.device("cuda:0")


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for rep in range(4):
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    marks = [("start", ev())]
    yd = y0_pinned.to(dev, non_blocking=True)
    marks.append(("h2d", ev()))
    s = de.OdeSystem(de.rhs_plugins.lorenz, yd, t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"], ensemble=True)
    marks.append(("construct", ev()))
    w1 = time.perf_counter()
    s.integrate()
    marks.append(("integrate", ev()))
    w2 = time.perf_counter()
    out_y.copy_(s[-1].y, non_blocking=True)
    out_c[0].copy_(s.accepted, non_blocking=True)
    out_c[1].copy_(s.rejected, non_blocking=True)
    out_c[2].copy_(s.status, non_blocking=True)
    marks.append(("d2h", ev()))
    torch.cuda.synchronize()
    w3 = time.perf_counter()
    if rep:
        print("rep %d: " % rep + "  ".join("%s %.2f ms" % (marks[i][0], marks[i - 1][1].elapsed_time(marks[i][1])) for i in range(1, len(marks))) +
              "  | host: construct %.2f integrate %.2f total %.2f ms" % ((w1 - w0) * 1e3, (w2 - w1) * 1e3, (w3 - w0) * 1e3))
