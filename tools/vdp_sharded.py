#!/usr/bin/env python
"""BASELINE config 4: Van der Pol ensemble, 2^24 trajectories with per-trajectory mu in [0.1, 10], RK45 adaptive,
sharded round-robin over the ranks of `torchrun`, NCCL termination all-reduce + final gather.

    python tools/vdp_sharded.py [--n 16777216]                     (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/vdp_sharded.py
Strong scaling: the GLOBAL ensemble is fixed; rank r integrates trajectories r, r+W, ...  Prints one JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import desolver_b200 as de  # noqa: E402
from desolver_b200 import distributed as dd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 24)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", type=int, default=2048, help="trajectories compared with the CPU oracle on rank 0")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = de.benchmarks.VDP
y0, mu = de.benchmarks.vdp_ensemble(args.n)
y0_shard = torch.as_tensor(np.ascontiguousarray(y0[:, rank::world])).to(dev)
mu_shard = torch.as_tensor(np.ascontiguousarray(mu[rank::world])).to(dev)


def make(y, consts):
    a = de.OdeSystem(de.rhs_plugins.van_der_pol, y, t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"],
                     constants=consts, ensemble=True)
    a.method = "RK45"
    return a


sh = dd.ShardedEnsemble.__new__(dd.ShardedEnsemble)        # shards were cut on the host: build the object directly
sh.group, sh.rank, sh.world_size, sh.n_global = None, rank, world, args.n
sh.system = make(y0_shard, dict(mu=mu_shard))
sh.statistics = None
times = []
for rep in range(args.reps):
    sh.system.reset()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stats = sh.integrate()                                    # local fused kernel + NCCL statistics all-reduce
    e1.record()
    torch.cuda.synchronize()
    t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    times.append(float(t_ms))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
full = sh.gather()
e1.record()
torch.cuda.synchronize()
gather_ms = e0.elapsed_time(e1)
local_att = float((sh.system.accepted + sh.system.rejected).sum())
if rank == 0:
    best = min(times)
    out = dict(config="van_der_pol n=%d RK45CK mu in [0.1,10] t in [0,10]" % args.n, n_gpus=world, ms=best, ms_all=times,
               trajectory_steps=stats["attempts"], trajectory_steps_per_s=stats["attempts"] / best * 1e3,
               algorithmic_GBps=stats["attempts"] * 72.0 / best / 1e6, hbm_roofline_frac_per_gpu=stats["attempts"] * 72.0 / best / 1e6 / world / 6538.3,
               failed=stats["failed"], gather_ms=gather_ms, rank0_local_attempts=local_att,
               load_imbalance=local_att * world / stats["attempts"])
    if args.check:
        from desolver_b200.integrators import tableaux as tb
        from oracle import oracle as orc
        m = args.check
        ref = orc.erk_ensemble(orc.RHS_VDP, y0[:, :m], cfg["t0"], cfg["tf"], cfg["dt0"], cfg["rtol"], cfg["atol"],
                               *tb.rk_arrays("RK45CKSolver"), params=mu[None, :m], threads=os.cpu_count())
        got = full["y"][:, :m].cpu().numpy()
        err = np.max(np.abs(got - ref["y"]), axis=0) / np.max(np.abs(ref["y"]), axis=0)
        same = (full["accepted"][:m].cpu().numpy() == ref["accepted"]) & (full["rejected"][:m].cpu().numpy() == ref["rejected"])
        out.update(checked=m, max_rel_err=float(err.max()), counts_identical=float(same.mean()))
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
