#!/usr/bin/env python
"""Correctness of the sharded path on N GPUs of one box (run under torchrun): the collectives behind the C ABI
(dsb_comm_allreduce_counters, dsb_comm_gather) and the scatter-store through CUDA-IPC peer buffers must reproduce the
unsharded single-GPU run bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multigpu_check.py
Prints one JSON line on rank 0 and exits non-zero on any mismatch."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import desolver_b200 as de  # noqa: E402
from desolver_b200 import distributed as dd  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = dict(world=world)
ok = True
for n in (100003, 1 << 18):                                   # a length that does not divide by the world size, and one that does
    cfg = de.benchmarks.VDP
    y0, mu = de.benchmarks.vdp_ensemble(n)
    y0_t, mu_t = torch.as_tensor(y0), torch.as_tensor(mu)

    def make(y, consts):
        s = de.OdeSystem(de.rhs_plugins.van_der_pol, y.to(dev), t=(cfg["t0"], 3.0), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"],
                         constants={k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in consts.items()}, ensemble=True,
                         device=dev)
        s.method = "RK45"
        return s
    ref = None
    if rank == 0:
        ref = make(y0_t, dict(mu=mu_t))
        ref.integrate()
    for scatter in (False, True, "push"):
        sh = dd.ShardedEnsemble(make, y0_t, dict(mu=mu_t), scatter_store=scatter is True, push=scatter == "push")
        stats = sh.integrate()
        full = sh.gather()
        torch.cuda.synchronize()
        if rank == 0:
            same = (torch.equal(full["y"], ref[-1].y) and torch.equal(full["t"], ref[-1].t) and torch.equal(full["accepted"], ref.accepted)
                    and torch.equal(full["rejected"], ref.rejected) and torch.equal(full["status"], ref.status)
                    and stats["attempts"] == ref.stats["attempts"] and stats["accepted"] == ref.stats["accepted"])
            out["n=%d scatter=%s" % (n, scatter)] = bool(same)
            ok = ok and same
        elif scatter is False:
            # dsb_comm_gather delivers the full ensemble on every rank
            chk = torch.tensor([float(full["y"].sum()), float(full["accepted"].sum())], dtype=torch.float64, device=dev)
        dist.barrier()
        del sh
if rank == 0:
    out["ok"] = bool(ok)
    print(json.dumps(out))
dist.barrier()
dd.destroy_communicators()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
