#!/usr/bin/env python
"""Times the UNMODIFIED reference's numpy backend on the headline workload:

  CPU-A  same semantics as the ensemble engine: one `desolver.OdeSystem` per trajectory (own dt, own error norm), a
         process per core, throughput = sum of step attempts / wall;
  CPU-B  the reference's best case: each process integrates its shard as ONE (3, N/P) state -- one global dt and one
         global error norm per shard, i.e. DIFFERENT semantics (every trajectory takes the steps the stiffest one needs).

    python tools/reference_numpy_timing.py [--tree /root/reference] [--seconds 20] > profiles/round2_reference_numpy_timing.json

TEST / MEASUREMENT INFRASTRUCTURE: needs a reference source tree (the build container's /root/reference, or
baseline/_ref on a box where the driver installed it); `autoray`, which the offline image lacks, is replaced by the
dispatch-only stand-in oracle/refshim (no arithmetic)."""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

os.environ.setdefault("OMP_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _paths(tree):
    # worker processes only: the reference prints a banner on import; stdout belongs to the caller's JSON line
    sys.stdout.flush()
    os.dup2(2, 1)
    try:
        import autoray  # noqa: F401
        extra = []
    except Exception:
        extra = [os.path.join(ROOT, "oracle", "refshim")]
    for p in extra + [tree, ROOT]:
        if p not in sys.path:
            sys.path.insert(0, p)


def _lorenz(t, y, sigma, rho, beta):
    import numpy as np
    return np.stack([sigma * (y[1] - y[0]), y[0] * (rho - y[2]) - y[1], y[0] * y[1] - beta * y[2]])


def _cpu_a(job):
    tree, lo, hi, budget = job
    _paths(tree)
    import numpy as np
    import desolver as de
    from desolver_b200 import benchmarks as bm
    cfg = bm.LORENZ
    y0 = bm.lorenz_ensemble(hi)[:, lo:hi]
    attempts, done, t0 = 0, 0, time.perf_counter()
    detail = []                                 # (global index, accepted, attempts, final state): bench.py's live parity check
    for i in range(y0.shape[1]):
        a = de.OdeSystem(_lorenz, y0=y0[:, i].copy(), t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"],
                         constants=dict(sigma=cfg["sigma"], rho=cfg["rho"], beta=cfg["beta"]))
        a.method = "RK45"
        calls = [0]
        inner = a.integrator.step

        def counted(*p, _inner=inner, **k):
            calls[0] += 1
            return _inner(*p, **k)
        a.integrator.step = counted
        a.integrate()
        attempts += calls[0]
        done += 1
        detail.append((lo + i, len(a) - 1, calls[0], np.asarray(a.y[-1], dtype=np.float64).copy()))
        if time.perf_counter() - t0 > budget:
            break
    return attempts, done, time.perf_counter() - t0, detail


def _cpu_b(job):
    tree, lo, hi, _ = job
    _paths(tree)
    import numpy as np
    import desolver as de
    from desolver_b200 import benchmarks as bm
    cfg = bm.LORENZ
    y0 = bm.lorenz_ensemble(hi)[:, lo:hi]
    t0 = time.perf_counter()
    a = de.OdeSystem(_lorenz, y0=np.ascontiguousarray(y0), t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"],
                     constants=dict(sigma=cfg["sigma"], rho=cfg["rho"], beta=cfg["beta"]))
    a.method = "RK45"
    calls = [0]
    inner = a.integrator.step

    def counted(*p, **k):
        calls[0] += 1
        return inner(*p, **k)
    a.integrator.step = counted
    a.integrate()
    return calls[0] * (hi - lo), hi - lo, time.perf_counter() - t0


def measure(tree, seconds=20.0, procs=None, details=None):
    """`details`: a list that receives CPU-A's per-trajectory results (global index, accepted steps, step attempts, final
    state) -- what the REFERENCE ITSELF computes on this box, for a live parity check of the GPU run."""
    procs = procs or max(1, len(os.sched_getaffinity(0)))
    out = dict(tree=tree, cores=procs, workload="lorenz63_rk45ck, trajectories of the benchmark ensemble (prefix), t in [0,2], tol 1e-9",
               unit="trajectory-steps/s")
    with ProcessPoolExecutor(procs) as pool:
        per = 4096                                           # upper bound per process; the time budget cuts it short
        jobs = [(tree, p * per, (p + 1) * per, seconds * 0.5) for p in range(procs)]
        t0 = time.perf_counter()
        res = list(pool.map(_cpu_a, jobs))
        wall = max(r[2] for r in res)                        # the slowest worker's integration loop: imports are not the reference's work
        if details is not None:
            for r in res:
                details.extend(r[3])
        out["cpu_a"] = dict(what="one reference OdeSystem per trajectory (same semantics as the engine)", value=sum(r[0] for r in res) / wall,
                            per_core=sum(r[0] for r in res) / wall / procs, trajectories=sum(r[1] for r in res), wall_s=wall)
        shard = 16384
        jobs = [(tree, p * shard, (p + 1) * shard, 0) for p in range(procs)]
        t0 = time.perf_counter()
        res = list(pool.map(_cpu_b, jobs))
        wall = max(r[2] for r in res)
        out["cpu_b"] = dict(what="each process integrates its shard as ONE (3, %d) state: global dt and norm per shard, different "
                                 "semantics" % shard, value=sum(r[0] for r in res) / wall, per_core=sum(r[0] for r in res) / wall / procs,
                            trajectories=sum(r[1] for r in res), wall_s=wall)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tree", default="/root/reference")
    ap.add_argument("--seconds", type=float, default=20.0)
    ap.add_argument("--details-out", default=None, help=".npz that receives CPU-A's per-trajectory results")
    args = ap.parse_args()
    det = [] if args.details_out else None
    rec = measure(args.tree, args.seconds, details=det)
    if det is not None:
        import numpy as np
        np.savez(args.details_out, index=np.array([d[0] for d in det]), accepted=np.array([d[1] for d in det]),
                 attempts=np.array([d[2] for d in det]), y=np.stack([d[3] for d in det], axis=1))
    print(json.dumps(rec, indent=1))
