"""The BASELINE.json configurations other than the headline, each as a function that bench.py calls and times under the
driver's clock: one JSON-able dict per config with its time, its own bound and a sampled parity check against the CPU
oracle (test infrastructure, used here only as the checker).

  config 2  linear y' = A y, A 4096 x 4096, 8192 columns, DP 8(7): FP64 tensor pipe (DMMA), 37.2 TFLOP/s at 1965 MHz
  config 3  N-body 64 bodies x 65536 systems, BABs9o7H, dt = 0.01, 10 steps: FP64 CUDA-core pipe, 34.2 TFLOP/s measured
            (tools/fp64_peak.cu), 16 FP64 instructions per pair interaction
  config 4  Van der Pol 2^24 trajectories, mu in [0.1, 10], RK45CK: 72 algorithmic bytes per trajectory-step
and, for N > 1 ranks, the two sharded runs VERDICT.md asked for: the 2^20 Lorenz ensemble split over the ranks (strong
scaling) and config 4 split over the ranks including the gather of the results.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DMMA_PEAK_TFLOPS = 37.2          # 148 SMs x 64 FMA/clk x 2 x 1.965 GHz
FP64_PEAK_TFLOPS = 34.2          # tools/fp64_peak.cu on this pool (DFMA issue-limited)


def _timed(torch, fn, reps, sync=None):
    out = []
    for _ in range(reps):
        if sync is not None:
            sync()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return out


def _threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def run_linear(dev, reps=2, check=8, n=4096, columns=8192, gemm="auto"):
    """Config 2.  gemm = "dmma": every product is dsb_dgemm (FP64 DMMA); "int8": the digit-plane product on the INT8
    tensor cores (dsb_ozaki_*); "auto" (the plugin's default): INT8 for the full-width products, dsb_dgemm for the
    compacted tail."""
    import torch
    import desolver_b200 as de
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    lin = de.rhs_plugins.linear
    previous, lin.gemm = lin.gemm, gemm
    try:
        return _run_linear(torch, de, tb, orc, dev, reps, check, n, columns, gemm)
    finally:
        lin.gemm = previous
        lin._int8.clear()


def _run_linear(torch, de, tb, orc, dev, reps, check, n, columns, gemm):
    lin = de.rhs_plugins.linear
    calls0 = (lin.calls_dsb, lin.calls_library, lin.calls_int8)
    cfg = de.benchmarks.LINEAR
    A, Y0 = de.benchmarks.linear_system(n, columns)
    Ad = torch.as_tensor(A).to(dev)
    a = de.OdeSystem(de.rhs_plugins.linear, torch.as_tensor(Y0).to(dev), t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"],
                     atol=cfg["atol"], constants=dict(A=Ad), ensemble=True, device=dev)
    a.method = "RK87"
    a.integrate()                                              # warm-up: allocator, graph capture
    ms_all = _timed(torch, lambda: (a.reset(), a.integrate()), reps)
    ms = min(ms_all)
    att = a.accepted + a.rejected
    flops_alg = float(att.sum()) * 14 * 2.0 * n ** 2           # (S + 1) * 2 n^2 per column-step attempt
    out = dict(workload="linear_4096x8192_rk8713m", gemm=gemm, ms=ms, ms_all=ms_all, column_step_attempts=int(att.sum()),
               attempts_per_column_max=int(att.max()), gemm_launches_dsb=int(lin.calls_dsb - calls0[0]),
               gemm_launches_library=int(lin.calls_library - calls0[1]), gemm_launches_int8=int(lin.calls_int8 - calls0[2]),
               tflops_algorithmic=flops_alg / ms / 1e9, bound="tensor (FP64 DMMA)", peak_tflops=DMMA_PEAK_TFLOPS,
               frac=flops_alg / ms / 1e9 / DMMA_PEAK_TFLOPS)
    if lin.calls_int8 - calls0[2] > 0:
        # 36 INT8 plane products per FP64 product; 4.5 POP/s is the dense INT8 tensor peak of the part (not measured here)
        out.update(bound="tensor (INT8 tcgen05, 36 plane products per FP64 product)", peak_tflops=None, frac=None,
                   int8_tops=36 * flops_alg / ms / 1e9, fp64_equivalent_tflops=flops_alg / ms / 1e9)
    if check:
        idx = np.linspace(0, columns - 1, check).astype(int)
        t0 = time.time()
        ref = orc.erk_ensemble(orc.RHS_LINEAR, Y0[:, idx], cfg["t0"], cfg["tf"], cfg["dt0"], cfg["rtol"], cfg["atol"],
                               *tb.rk_arrays("RK8713MSolver"), shared=A, threads=_threads())
        got = a[-1].y[:, idx].cpu().numpy()
        err = np.max(np.abs(got - ref["y"]), axis=0) / np.max(np.abs(ref["y"]), axis=0)
        same = (a.accepted[idx].cpu().numpy() == ref["accepted"]) & (a.rejected[idx].cpu().numpy() == ref["rejected"])
        out["parity"] = dict(checked_columns=int(check), counts_identical_frac=float(same.mean()), max_rel_err=float(err.max()),
                             oracle_seconds=time.time() - t0)
    return out


def run_nbody(dev, reps=3, check=16, systems=65536, bodies=64):
    import torch
    import desolver_b200 as de
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    state, m = de.benchmarks.nbody_ensemble(systems, bodies=bodies, seed=0)
    y0 = torch.as_tensor(np.ascontiguousarray(np.moveaxis(state, 1, -1))).to(dev)
    a = de.OdeSystem(de.rhs_plugins.nbody, y0, t=(0.0, 0.1), dt=0.01, constants=dict(m=torch.as_tensor(m), eps2=1e-4, G=1.0),
                     ensemble=True, device=dev)
    a.method = "BABS9O7H"
    a.integrate()
    ms_all = _timed(torch, lambda: (a.reset(), a.integrate()), reps)
    ms = min(ms_all)
    steps = int(a.accepted.sum())
    kicks = int((a.integrator.tableau_intermediate[:, 2] != 0).sum())
    pairs = steps * kicks * bodies * bodies
    fp64_tflops = pairs * 16 * 2.0 / ms / 1e9                   # 16 FP64 instructions per pair, counted as FMA (2 flop)
    out = dict(workload="nbody_64x65536_babs9o7h_10steps", ms=ms, ms_all=ms_all, system_steps_per_s=steps / ms * 1e3,
               pair_interactions_per_s=pairs / ms * 1e3, bound="fp64 pipe", fp64_instr_per_pair=16,
               tflops_fp64_issue=fp64_tflops, peak_tflops=FP64_PEAK_TFLOPS, frac=fp64_tflops / FP64_PEAK_TFLOPS)
    if check:
        idx = np.linspace(0, systems - 1, check).astype(int)
        sub = np.ascontiguousarray(np.moveaxis(state[:, idx], 1, -1).reshape(-1, check))
        shared = np.concatenate([[1.0, 1e-4], m])
        ref = orc.symplectic_ensemble(orc.RHS_NBODY, sub, 0.0, 0.1, 0.01, tb.symplectic_array("BABs9o7HSolver")[0], shared=shared,
                                      threads=_threads())
        got = a[-1].y.reshape(-1, systems)[:, idx].cpu().numpy()
        err = np.max(np.abs(got - ref["y"]), axis=0) / np.max(np.abs(ref["y"]), axis=0)
        out["parity"] = dict(checked_systems=int(check), steps_identical=bool(np.all(a.accepted[idx].cpu().numpy() == ref["steps"])),
                             max_rel_err=float(err.max()))
    return out


def _vdp_parity(full, y0, mu, cfg, check):
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    m = check
    ref = orc.erk_ensemble(orc.RHS_VDP, np.ascontiguousarray(y0[:, :m]), cfg["t0"], cfg["tf"], cfg["dt0"], cfg["rtol"], cfg["atol"],
                           *tb.rk_arrays("RK45CKSolver"), params=mu[None, :m], threads=_threads())
    got = full["y"][:, :m].cpu().numpy()
    err = np.max(np.abs(got - ref["y"]), axis=0) / np.max(np.abs(ref["y"]), axis=0)
    same = (full["accepted"][:m].cpu().numpy() == ref["accepted"]) & (full["rejected"][:m].cpu().numpy() == ref["rejected"])
    return dict(checked_trajectories=int(m), counts_identical_frac=float(same.mean()), max_rel_err=float(err.max()))


def run_vdp(dev, world=1, rank=0, reps=3, check=2048, n=1 << 24, hbm_peak=6538.3, scatter=True):
    """Config 4.  world == 1: the whole ensemble on this GPU.  world > 1: round-robin shards, the statistics
    all-reduce enqueued behind each rank's kernel, and the results gathered -- three ways, all timed: "push" (chunks of
    finished trajectories copied to rank 0's staging buffer over NVLink while the shard is still integrating, one
    interleave kernel on rank 0), "scatter" (the kernel's own stores go to rank 0's buffer), "gather" (dsb_comm_gather
    afterwards: ncclAllGather + interleave on every rank).  Times are max over ranks; `ms` INCLUDES the gather."""
    import torch
    import torch.distributed as dist
    import desolver_b200 as de
    from desolver_b200 import distributed as dd
    cfg = de.benchmarks.VDP
    y0, mu = de.benchmarks.vdp_ensemble(n)

    def make(y, consts):
        s = de.OdeSystem(de.rhs_plugins.van_der_pol, y, t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"],
                         constants=consts, ensemble=True, device=dev)
        s.method = "RK45"
        return s
    y0_shard = torch.as_tensor(np.ascontiguousarray(y0[:, rank::world])).to(dev)
    mu_shard = torch.as_tensor(np.ascontiguousarray(mu[rank::world])).to(dev)

    def build(mode):
        sh = dd.ShardedEnsemble.__new__(dd.ShardedEnsemble)     # shards were cut on the host: build the object directly
        sh.group, sh.rank, sh.world_size, sh.n_global, sh.root = None, rank, world, n, 0
        sh.system = make(y0_shard, dict(mu=mu_shard))
        sh._global = sh._stats = sh._err = sh._side = sh.scatter = sh.push = None
        s = sh.system
        if mode == "scatter" and world > 1:
            sh.scatter = dd.ScatterTarget(s.numel, n, s.device, None, 0)
            s._scatter_store = dict(index_mul=world, index_add=rank, stride=n,
                                    **{k: sh.scatter.pointer(k) for k in ("y", "t", "dt", "accepted", "rejected", "status")})
        if mode == "push" and world > 1:
            sh.push = dd.PushTarget(s.numel, n, s.device, None, 0)
            s._push_target = dict(n_max=sh.push.n_max, **{k: sh.push.slot(k) for k, _ in dd.PushTarget.FIELDS})
        return sh

    def barrier():
        if world > 1:
            dist.barrier()

    results = {}
    for mode in (["push", "scatter", "gather"] if (world > 1 and scatter) else ["gather"]):
        sh = build(mode)
        full = [None]
        stats = [None]

        def one():
            sh.system.reset()
            sh.integrate(wait=False)                            # kernel + all-reduce enqueued, nothing read back
            full[0] = sh.gather() if world > 1 else None
            stats[0] = sh.statistics

        one()                                                   # warm-up (communicator, IPC mapping, allocator)
        ms_all = _timed(torch, one, reps, sync=barrier)
        t = torch.tensor([min(ms_all)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        results[mode] = (float(t[0]), ms_all, stats[0], full[0], sh)
    best_mode = min(results, key=lambda k: results[k][0])
    ms, ms_all, stats, full, sh = results[best_mode]
    att = stats["attempts"]
    out = dict(workload="vdp_2^24_rk45ck", n_gpus=world, ms=ms, ms_all_rank0=ms_all, mode=best_mode if world > 1 else "single",
               includes_gather=world > 1, trajectory_steps=att, trajectory_steps_per_s=att / ms * 1e3,
               bound="hbm (72 algorithmic bytes per trajectory-step)", algorithmic_GBps=att * 72.0 / ms / 1e6,
               peak_GBps=hbm_peak * world, frac=att * 72.0 / ms / 1e6 / (hbm_peak * world), failed=stats["failed"])
    for k, v in results.items():
        out["ms_" + k] = v[0]
    if check and rank == 0:
        if full is None:
            s = sh.system
            full = dict(y=s[-1].y, accepted=s.accepted, rejected=s.rejected)
        out["parity"] = _vdp_parity(full, y0, mu, cfg, check)
    return out


def run_lorenz_strong(dev, world, rank, reps=5, n=1 << 20):
    """The 2^20 Lorenz ensemble of the headline split over the ranks (strong scaling): n / world trajectories per GPU,
    statistics all-reduce behind each kernel, max over ranks."""
    import torch
    import torch.distributed as dist
    import desolver_b200 as de
    from desolver_b200 import distributed as dd
    cfg = de.benchmarks.LORENZ
    y0 = de.benchmarks.lorenz_ensemble(n)[:, rank::world]
    sh = dd.ShardedEnsemble.__new__(dd.ShardedEnsemble)
    sh.group, sh.rank, sh.world_size, sh.n_global, sh.root = None, rank, world, n, 0
    s = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(np.ascontiguousarray(y0)).to(dev), t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"],
                     rtol=cfg["rtol"], atol=cfg["atol"], ensemble=True, device=dev)
    s.method = "RK45"
    sh.system = s
    sh._global = sh._stats = sh._err = sh._side = sh.scatter = None
    stats = [None]

    def one():
        s.reset()
        sh.integrate(wait=False)
        stats[0] = sh.statistics

    one()
    ms_all = _timed(torch, one, reps, sync=(dist.barrier if world > 1 else None))
    t = torch.tensor([min(ms_all)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    att = stats[0]["attempts"]
    return dict(workload="lorenz63_rk45ck_2^20_total", n_gpus=world, trajectories_per_gpu=n // world, ms=ms, ms_all_rank0=ms_all,
                trajectory_steps=att, trajectory_steps_per_s=att / ms * 1e3, scaling="strong")
