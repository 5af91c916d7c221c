#!/usr/bin/env python
"""Where the end-to-end pass (pinned host y0 -> pinned host result through OdeSystem(host_io=True)) spends its time:
device-resident pass, host pipeline with a reused system, with construction, for several chunk counts; host-side clock of
the constructor and of integrate()."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import desolver_b200 as de  # noqa: E402

dev = torch.device("cuda:0")
cfg = de.benchmarks.LORENZ
n = 1 << 20
y0 = de.benchmarks.lorenz_ensemble(n, seed=0)
y0_pinned = torch.as_tensor(y0).pin_memory()
host_out = dict(y=torch.empty((3, n), dtype=torch.float64).pin_memory())
kw = dict(t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"], ensemble=True, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=6, warm=2):
    out = []
    for i in range(warm + reps):
        flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= warm:
            out.append(a.elapsed_time(b))
    return float(np.median(out)), out


res = {}
s = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0).to(dev), **kw)
s.method = "RK45CK"
res["device_resident_ms"] = timed(lambda: (s.reset(), s.integrate()))[0]
for chunks in (8, 16, 32, 64):
    holder = {}

    def build():
        holder["s"] = None
        holder["s"] = de.OdeSystem(de.rhs_plugins.lorenz, y0_pinned, host_io=True, host_out=host_out, host_chunks=chunks, **kw)
        holder["s"].integrate()
    res["e2e_construct_chunks%d_ms" % chunks] = timed(build)[0]
    s2 = holder["s"]
    res["e2e_reuse_chunks%d_ms" % chunks] = timed(lambda: (s2.reset(), s2.integrate()))[0]
# host-side clock of one pass (16 chunks)
torch.cuda.synchronize()
t0 = time.perf_counter()
s3 = de.OdeSystem(de.rhs_plugins.lorenz, y0_pinned, host_io=True, host_out=host_out, host_chunks=16, **kw)
t1 = time.perf_counter()
s3.integrate()
t2 = time.perf_counter()
torch.cuda.synchronize()
t3 = time.perf_counter()
res["host_clock_ms"] = dict(constructor=(t1 - t0) * 1e3, integrate_call=(t2 - t1) * 1e3, sync_after=(t3 - t2) * 1e3)
print(json.dumps(res, indent=1))
