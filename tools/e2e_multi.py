#!/usr/bin/env python
"""e2e (host_io) pass time per rank for several chunk counts, under torchrun: where does the N-GPU e2e efficiency go?"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import desolver_b200 as de  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 20
cfg = de.benchmarks.LORENZ
y0 = torch.from_numpy(np.ascontiguousarray(de.benchmarks.lorenz_ensemble(n * world)[:, rank::world])).pin_memory()
host_out = dict(y=torch.empty((3, n), dtype=torch.float64).pin_memory(), accepted=torch.empty(n, dtype=torch.int32).pin_memory(),
                rejected=torch.empty(n, dtype=torch.int32).pin_memory())
res = {}
for chunks in (4, 8, 16, 32, 64):
    times = []
    s = None
    for i in range(7):
        s = None
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"], ensemble=True,
                         device=dev, host_io=True, host_out=host_out, host_chunks=chunks)
        s.integrate()
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            times.append(a.elapsed_time(b))
    t = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
    allt = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(allt, t)
    else:
        allt = [t]
    res[chunks] = [round(float(x), 3) for x in allt]
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
