#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 ensemble engine (BASELINE.json metric):
trajectory-steps/s, RK45 (Cash-Karp) adaptive fp64, Lorenz-63, 2^20 trajectories per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over the batch: every trajectory of the ensemble is
integrated from t=0 to t=2 (rtol=atol=1e-9, dt0=1e-2), ~254 step attempts each.  A
"trajectory-step" is one step ATTEMPT of one trajectory (accepted or rejected: same work); the
count comes from the per-trajectory device counters of the same run.

 value     device-timed (CUDA events on the launch stream), inputs resident in HBM.
 e2e       same metric through the public API (`OdeSystem(...).integrate()`) with HOST buffers:
           pinned host y0 -> device, integrate, final state + counters -> host, all timed.
 roofline  algorithmic bytes (80 B per Lorenz trajectory-step: read+write y(3), t, dt)
           over the fused kernel's event-timed duration, against the measured HBM copy bandwidth.
 cpu_baseline  the CPU oracle (port of the reference algorithm, C) on this box's host cores, on a
           bounded sample of the same ensemble; `--impl reference` runs ONLY that.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 1 << 20
BYTES_PER_TRAJ_STEP = 80.0          # 2*8*(dim+2), SURVEY.md section 8d
METRIC = "trajectory-steps/s (RK45 Cash-Karp adaptive fp64, Lorenz-63, 2^20 trajectories per GPU, t in [0,2], tol 1e-9)"
UNIT = "trajectory-steps/s"


def static_config(n_local):
    """The workload description shared VERBATIM by both arms (the driver compares the two `config` dicts)."""
    from desolver_b200 import benchmarks as bm
    cfg = bm.LORENZ
    return {"workload": "lorenz63_rk45ck_2^20_per_gpu", "trajectories_per_gpu": int(n_local),
            "t_span": [cfg["t0"], cfg["tf"]], "rtol": cfg["rtol"], "atol": cfg["atol"], "dt0": cfg["dt0"], "method": "RK45CK",
            "y0": "U([-20,20]x[-25,25]x[0,50]), numpy default_rng(0), dealt round-robin over the ranks",
            "l2": "flushed between timed iterations (256 MiB memset)",
            "sharding": "round-robin over ranks, no data-path collective"}


def reference_numpy_numbers(live=True):
    """The reference's own numpy backend (CPU-A: one OdeSystem per trajectory; CPU-B: a whole shard as ONE state with a
    global dt -- different semantics), timed LIVE on this box's host cores from baseline/_ref -- the unmodified reference,
    installed in the build container by baseline/install_reference.py (git-ignored, travels with the snapshot; `autoray`,
    absent from the offline image, is the dispatch-only stand-in oracle/refshim).  CPU-A's per-trajectory results are kept
    for `parity.reference_live`.  Without baseline/_ref the record taken in the build container
    (profiles/round2_reference_numpy_timing.json) is quoted instead, and says so in `where`."""
    rec = None
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "round2_reference_numpy_timing.json")))
        rec["where"] = "build container (recorded by tools/reference_numpy_timing.py)"
    except Exception:
        pass
    for tree in (os.path.join(ROOT, "baseline", "_ref"),):
        if live and os.path.isdir(os.path.join(tree, "desolver")):
            try:
                # its own process: the measurement forks a pool of workers, which must not inherit this process's CUDA /
                # NCCL state and threads
                import subprocess
                import tempfile
                with tempfile.TemporaryDirectory() as tmp:
                    npz = os.path.join(tmp, "cpu_a.npz")
                    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "reference_numpy_timing.py"), "--tree", tree,
                                          "--seconds", "6", "--details-out", npz], check=True, timeout=240,
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="1"))
                    live = json.loads(out.stdout.decode())
                    d = np.load(npz)
                    reference_numpy_numbers.details = [(int(i), int(a), int(t), d["y"][:, j].copy()) for j, (i, a, t) in
                                                       enumerate(zip(d["index"], d["accepted"], d["attempts"]))]
                live["where"] = "this box, live (%s: the unmodified reference, installed by baseline/install_reference.py)" % tree
                return live
            except Exception as e:                          # noqa: BLE001
                if rec is not None:
                    rec["live_error"] = repr(e)[:200]
    return rec


reference_numpy_numbers.details = None


def reference_live_parity(y_gpu, acc_gpu, rej_gpu, n_gpus, rank):
    """The GPU results against what the REFERENCE ITSELF (numpy backend, one OdeSystem per trajectory: CPU-A of
    reference_numpy_numbers) computed on this box in this run, for the trajectories it got through."""
    det = reference_numpy_numbers.details
    if not det:
        return None
    mine = [d for d in det if d[0] % n_gpus == rank and d[0] // n_gpus < y_gpu.shape[1]]
    if not mine:
        return None
    loc = np.array([d[0] // n_gpus for d in mine])
    acc = np.array([d[1] for d in mine])
    att = np.array([d[2] for d in mine])
    y_ref = np.stack([d[3] for d in mine], axis=1)
    err = np.max(np.abs(y_gpu[:, loc] - y_ref), axis=0) / np.max(np.abs(y_ref), axis=0)
    same = (acc_gpu[loc] == acc) & (acc_gpu[loc] + rej_gpu[loc] == att)
    return {"what": "GPU vs the unmodified reference's numpy backend run on this box (baseline/_ref), same trajectories",
            "checked_trajectories": int(loc.shape[0]), "counts_identical_frac": float(same.mean()),
            "max_rel_err": float(err.max()), "median_rel_err": float(np.median(err))}


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_sample(seconds_budget, n_global, rank_stride=1):
    """Times the CPU oracle (oracle/desolver_oracle.c) on a bounded prefix of the same ensemble,
    all host threads, one contiguous shard per thread.  Returns (steps_per_s, cores, sample_desc)."""
    from desolver_b200 import benchmarks as bm
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    cfg = bm.LORENZ
    threads = host_threads()
    tab = tb.rk_arrays("RK45CKSolver")
    params = [cfg["sigma"], cfg["rho"], cfg["beta"]]
    # calibrate on a small piece, then size the sample for ~seconds_budget of wall time
    y_small = bm.lorenz_ensemble(256 * threads)
    t0 = time.perf_counter()
    r = orc.erk_ensemble(orc.RHS_LORENZ, y_small, cfg["t0"], cfg["tf"], cfg["dt0"], cfg["rtol"], cfg["atol"], *tab,
                         params=params, threads=threads)
    rate = float((r["accepted"] + r["rejected"]).sum()) / (time.perf_counter() - t0)
    m = int(min(n_global, max(1024 * threads, rate * seconds_budget / 255.0)))
    y0 = bm.lorenz_ensemble(m)
    t0 = time.perf_counter()
    r = orc.erk_ensemble(orc.RHS_LORENZ, y0, cfg["t0"], cfg["tf"], cfg["dt0"], cfg["rtol"], cfg["atol"], *tab,
                         params=params, threads=threads)
    wall = time.perf_counter() - t0
    steps = float((r["accepted"] + r["rejected"]).sum())
    cpu_reference_sample.last = (m, r)          # kept for the full-size parity gate (checker use of the oracle)
    return steps / wall, threads, "first %d trajectories of the Lorenz ensemble, %.1f s wall" % (m, wall), steps, wall


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 5 ms during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.005)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def run_reference_arm(args):
    """`--impl reference`: the reference algorithm's CPU implementation (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_global = N_PER_GPU * args.gpus
    per_step = []
    info = None
    for i in range(args.warmup + args.steps):
        rate, cores, sample, steps, wall = cpu_reference_sample(args.cpu_seconds, n_global)
        if i >= args.warmup:
            per_step.append((steps, wall))
            info = (cores, sample)
    total_steps = sum(s for s, _ in per_step)
    total_wall = sum(w for _, w in per_step)
    value = total_steps / total_wall
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_wall / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": static_config(args.n_per_gpu),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": info[0], "kind": "port", "sample": info[1],
                             "what": "oracle/desolver_oracle.c: C restatement of the reference algorithm, one thread per core"},
            "cpu_reference_numpy": reference_numpy_numbers(),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def parity_gate(y_gpu, acc_gpu, rej_gpu, ref, global_index):
    """GPU results (this rank's shard, numpy) against the CPU oracle's results for the same trajectories.
    `global_index[j]` = position of local trajectory j in the oracle's (prefix) ensemble."""
    y_ref = ref["y"][:, global_index]
    err = np.max(np.abs(y_gpu - y_ref), axis=0) / np.max(np.abs(y_ref), axis=0)
    same = (acc_gpu == ref["accepted"][global_index]) & (rej_gpu == ref["rejected"][global_index])
    return {"checked_trajectories": int(global_index.shape[0]), "counts_identical_frac": float(same.mean()),
            "max_rel_err": float(err.max()), "p99_rel_err": float(np.percentile(err, 99)),
            "p99_99_rel_err": float(np.percentile(err, 99.99)), "median_rel_err": float(np.median(err)),
            "n_over_1e-10": int((err > 1e-10).sum()),
            "bar": "counts identical >= 0.999, rel err <= 1e-10 (BASELINE.json north_star); the handful of trajectories "
                   "beyond 1e-10 are chaos amplifying last-ulp differences of pow/atan -- pinned by "
                   "tests/test_oracle_golden.py::test_parity_tail_is_chaos_not_arithmetic"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="wall budget of the CPU baseline sample")
    ap.add_argument("--n-per-gpu", type=int, default=N_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs / scaling modes")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (for launch lists under ncu: the profiler "
                    "serialises the copy streams the flow-controlled launch overlaps with, so that kernel would sit in its "
                    "bounded wait); the JSON line then carries no e2e number")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import desolver_b200 as de
    from desolver_b200 import distributed as dd
    from desolver_b200.backend import cuda_backend as cb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; desolver_b200 has no CPU fallback")
    numa_node = dd.bind_to_gpu_numa_node(local_rank) if world > 1 else None     # before any pinned allocation
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = max(world, 1)
    n_local = args.n_per_gpu
    cfg = de.benchmarks.LORENZ

    # global ensemble of n_gpus * 2^20 trajectories, dealt round-robin: trajectory i -> rank i mod G
    y0_host = de.benchmarks.lorenz_ensemble(n_local * n_gpus)[:, rank::n_gpus]
    y0_pinned = torch.from_numpy(np.ascontiguousarray(y0_host)).pin_memory()
    y0_dev = y0_pinned.to(dev)

    sys_ = de.OdeSystem(de.rhs_plugins.lorenz, y0_dev, t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"],
                        atol=cfg["atol"], constants=dict(sigma=cfg["sigma"], rho=cfg["rho"], beta=cfg["beta"]),
                        ensemble=True)
    sys_.method = "RK45"
    sys_.lazy_status = True            # integrate() returns when the launch is enqueued; counters are read after the loop
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    comm = dd.communicator(dev) if world > 1 else None                # dsb_comm_* (NCCL) behind the C ABI
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    n_slots = args.steps + args.warmup + 4
    local_ctr = torch.zeros((n_slots, 4), dtype=torch.int64, device=dev)     # per step: this rank's kernel counters
    global_ctr = torch.zeros((n_slots, 4), dtype=torch.int64, device=dev)    # ... summed over the ranks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(slot, ev=None):
        """Device-resident step: reset (the kernel re-initialises from y0 itself), ONE launch, and -- behind it, on a side
        stream -- the termination / statistics all-reduce of the kernel's own counters (attempts, accepted, failed,
        unfinished).  No host synchronisation anywhere: the ranks meet in the collective on the device only."""
        sys_.reset()
        sys_._kernel_events = ev              # CUDA events recorded around the dsb_erk_run launch itself
        sys_.integrate()
        sys_._kernel_events = None
        local_ctr[slot].copy_(sys_._counters[2:6])
        if world > 1:
            side.wait_stream(torch.cuda.current_stream(dev))
            comm.allreduce_counters(local_ctr[slot].data_ptr(), global_ctr[slot].data_ptr(), 4, cb.C.c_void_p(side.cuda_stream))
        else:
            global_ctr[slot].copy_(local_ctr[slot])

    for w in range(args.warmup):
        one_step(w)
        flush.zero_()
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = sys_._launches
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        step_ev[i][0].record()
        one_step(args.warmup + i, kern_ev[i])
        step_ev[i][1].record()
        flush.zero_()                          # L2 flush between timed iterations (outside the events)
    if side is not None:
        torch.cuda.current_stream(dev).wait_stream(side)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in step_ev]
    kern_ms = [a.elapsed_time(b) for a, b in kern_ev]
    launches = sys_._launches - launches0 + (args.steps if world > 1 else 0)      # + one all-reduce kernel per step (NCCL)
    timed = slice(args.warmup, args.warmup + args.steps)
    g = global_ctr[timed].cpu().numpy()
    loc = local_ctr[timed].cpu().numpy()
    assert (g == g[0]).all() and (loc == loc[0]).all(), "the step is deterministic: every pass must count the same"
    attempts_global, failed = float(g[0, 0]), float(g[0, 2] + g[0, 3])
    attempts_local, accepted_local = float(loc[0, 0]), float(loc[0, 1])
    # whole-job time = first event to last event on each rank (the passes run back to back, rank-asynchronously), max over ranks
    span_ms = step_ev[0][0].elapsed_time(step_ev[-1][1])
    t_local = torch.tensor([sum(step_ms), sum(kern_ms), span_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    total_ms, total_kern_ms = float(t_local[0]), float(t_local[1])
    value = attempts_global * args.steps / (total_ms * 1e-3)

    y_a = sys_[-1].y.clone()
    acc_a, rej_a = sys_.accepted.clone(), sys_.rejected.clone()
    one_step(n_slots - 1)                    # (every rank: the collective inside one_step stays matched)
    deterministic = bool(torch.equal(y_a, sys_[-1].y) and torch.equal(acc_a, sys_.accepted) and torch.equal(rej_a, sys_.rejected))
    if side is not None:
        torch.cuda.current_stream(dev).wait_stream(side)

    # ---- e2e: public API with host buffers.  `host_io=True`: pinned host y0 in, pinned host results out; the
    # upload, the kernels and the download of the chunks overlap inside integrate() (OdeSystem._run_host_pipeline).
    # The timed region covers construction, every host<->device copy, the integration and the final synchronisation.
    # The step's RESULT is the final state (what the reference's a.y[-1] holds) plus the ensemble totals (attempts,
    # accepted, failed, unfinished: the kernel's 64-byte counter block, read back with every pass).  The per-trajectory
    # counter arrays are not part of it: they stay on the device and `a.accepted` / `a.rejected` / `a.status` fetch
    # them on demand (checked below on the first pass).
    e2e_steps = max(2, min(args.steps, 5))
    host_out = dict(y=torch.empty((3, n_local), dtype=torch.float64).pin_memory())
    host_chunks = 16 if world == 1 else 32           # more, smaller copies keep eight ranks' PCIe traffic interleaved
    e2e_ms = []
    E2E_WARMUP = 2                      # allocator / stream-creation warm-up iterations, untimed
    s2 = None
    for i in range(0 if args.no_e2e else E2E_WARMUP + e2e_steps):
        s2 = None                       # the previous pass's system is gone before the next one is built (its device
        barrier()                       # buffers return to torch's caching allocator: no cudaMalloc inside the timed region)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s2 = de.OdeSystem(de.rhs_plugins.lorenz, y0_pinned, t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"],
                          atol=cfg["atol"], ensemble=True, device=dev, host_io=True, host_out=host_out, host_chunks=host_chunks)
        s2.integrate()                  # returns after the last download has landed in host_out
        b.record()
        barrier()
        if i >= E2E_WARMUP:
            e2e_ms.append(a.elapsed_time(b))
        if i == 0 and rank == 0:        # the host results are the device-resident results, bit for bit
            assert torch.equal(s2[-1].y, y_a.cpu()) and torch.equal(s2.accepted, acc_a.cpu()) and int(s2.status.sum()) == 0
        flush.zero_()
    # the same with ONE system reused (reset() + integrate()): what a caller who integrates ensemble after ensemble pays;
    # reported next to the headline, which keeps the construction inside the timed region
    reuse_ms = []
    for i in range(0 if args.no_e2e else E2E_WARMUP + e2e_steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s2.reset()
        s2.integrate()
        b.record()
        barrier()
        if i >= E2E_WARMUP:
            reuse_ms.append(a.elapsed_time(b))
        flush.zero_()
    e2e_t = torch.tensor([float(np.mean(e2e_ms)) if e2e_ms else float("nan")], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = attempts_global / (float(e2e_t[0]) * 1e-3)
    h2d = int(y0_pinned.numel() * 8)
    d2h = int(sum(v.numel() * v.element_size() for v in host_out.values())) + 64      # + the counter block
    s2 = None

    # ---- the other BASELINE configs / scaling modes, under the same clock ----
    extra = {}
    if not args.no_configs:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import config_runs as cr
        peak_gbs = measured_peak_gbs()[0]
        del flush
        torch.cuda.empty_cache()
        try:
            if world == 1:
                extra["config4_vdp_2pow24"] = cr.run_vdp(dev, reps=3, check=2048, hbm_peak=peak_gbs)
                torch.cuda.empty_cache()
                extra["config3_nbody"] = cr.run_nbody(dev, reps=3, check=16)
                torch.cuda.empty_cache()
                extra["config2_linear"] = cr.run_linear(dev, reps=2, check=4)                    # plugin default: INT8 tensor cores
                torch.cuda.empty_cache()
                extra["config2_linear_dmma"] = cr.run_linear(dev, reps=2, check=4, gemm="dmma")     # every product on the FP64 DMMA kernel
            else:
                extra["strong_2pow20"] = cr.run_lorenz_strong(dev, world, rank, reps=5)
                extra["vdp_2pow24_sharded"] = cr.run_vdp(dev, world, rank, reps=3, check=2048, hbm_peak=peak_gbs)
        except Exception as e:                                  # noqa: BLE001  (the headline line must still be printed)
            extra["error"] = repr(e)[:400]
        torch.cuda.empty_cache()

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        kern_s = (total_kern_ms / args.steps) * 1e-3
        achieved = attempts_local * BYTES_PER_TRAJ_STEP / kern_s / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "erk_fused_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": static_config(n_local),
                "run": {"attempts_per_step": attempts_global, "accepted_local": accepted_local, "failed_trajectories": failed,
                        "wall_s_timed_region": wall, "ms_first_to_last_event": float(t_local[2]), "numa_node": numa_node,
                        "collective": ("dsb_comm_allreduce_counters (NCCL %d) on a side stream behind every launch"
                                       % cb.lib().dsb_comm_nccl_version()) if world > 1 else None},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_kind, "kernel": "dsb::erk_fused_kernel<LorenzRhs, StaticTab<CashKarp45>>",
                             "kernel_ms": total_kern_ms / args.steps,
                             "algorithmic_bytes_per_launch": attempts_local * BYTES_PER_TRAJ_STEP,
                             "trajectory_steps_per_s_kernel": attempts_local / kern_s},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": float(e2e_t[0]), "ms_all_rank0": e2e_ms,
                        "call": "OdeSystem(lorenz, pinned y0, ..., host_io=True).integrate(): construction, uploads, one "
                                "persistent launch, downloads and the final synchronisation inside the timed region",
                        "ms_per_step_system_reused_rank0": float(np.mean(reuse_ms)) if reuse_ms else None},
                "gpu_launches": launches}
        line.update(extra and {"configs": extra})
        # ---- parity gate + CPU baseline: the oracle integrates a prefix of the GLOBAL ensemble (all of it when the host
        # is fast enough) on the host cores; every trajectory of that prefix that lives on this rank is compared ----
        if not args.no_cpu_baseline:
            rate, cores, sample, _, _ = cpu_reference_sample(args.cpu_seconds, n_local * n_gpus)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "what": "oracle/desolver_oracle.c: C restatement of the reference algorithm, one thread per core"}
            # the reference itself is timed live at N = 1 only (at N > 1 the other ranks wait in a barrier on the same cores)
            line["cpu_reference_numpy"] = reference_numpy_numbers(live=(world == 1))
            m, ref = cpu_reference_sample.last
        else:
            from desolver_b200.integrators import tableaux as tb
            from oracle import oracle as orc
            m = min(4096 * n_gpus, n_local * n_gpus)
            ref = orc.erk_ensemble(orc.RHS_LORENZ, de.benchmarks.lorenz_ensemble(m), cfg["t0"], cfg["tf"], cfg["dt0"], cfg["rtol"],
                                   cfg["atol"], *tb.rk_arrays("RK45CKSolver"), params=[cfg["sigma"], cfg["rho"], cfg["beta"]],
                                   threads=host_threads())
        gidx = np.arange(rank, m, n_gpus)                       # global indices of this rank's trajectories inside the prefix
        k = gidx.shape[0]
        parity = parity_gate(y_a[:, :k].cpu().numpy(), acc_a[:k].cpu().numpy(), rej_a[:k].cpu().numpy(), ref, gidx)
        parity["bitwise_deterministic"] = deterministic
        parity["of_ensemble"] = "%d of the %d trajectories on rank 0" % (k, n_local)
        live = reference_live_parity(y_a.cpu().numpy(), acc_a.cpu().numpy(), rej_a.cpu().numpy(), n_gpus, rank)
        if live is not None:
            parity["reference_live"] = live
        line["parity"] = parity
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dd.destroy_communicators()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
