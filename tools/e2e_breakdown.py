#!/usr/bin/env python
"""Where the end-to-end (host buffers) time of one Lorenz 2^20 pass goes, for the host_io pipelines of OdeSystem:
"streamed" (one persistent launch + stream memory operations, dsb_erk_run_host) and "chunked" (one launch per chunk)."""
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
dev = torch.device("cuda:0")
host_out = dict(y=torch.empty((3, n), dtype=torch.float64).pin_memory(), accepted=torch.empty(n, dtype=torch.int32).pin_memory(),
                rejected=torch.empty(n, dtype=torch.int32).pin_memory(), status=torch.empty(n, dtype=torch.int32).pin_memory())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

for mode, chunks in (("streamed", 4), ("streamed", 8), ("streamed", 16), ("streamed", 32), ("chunked", 4), ("chunked", 8)):
    res = []
    for rep in range(6):
        flush.zero_()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        s = de.OdeSystem(de.rhs_plugins.lorenz, y0_pinned, t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"],
                         ensemble=True, device=dev, host_io=True, host_out=host_out, host_chunks=chunks)
        s.host_pipeline = mode
        b.record()
        w1 = time.perf_counter()
        s.integrate()
        c.record()
        torch.cuda.synchronize()
        w2 = time.perf_counter()
        if rep >= 2:
            res.append((a.elapsed_time(b), b.elapsed_time(c), (w1 - w0) * 1e3, (w2 - w1) * 1e3))
    r = np.mean(res, axis=0)
    print("%-8s chunks %2d: events construct %.3f integrate %.3f total %.3f ms | host construct %.3f integrate %.3f ms" %
          (mode, chunks, r[0], r[1], r[0] + r[1], r[2], r[3]))
