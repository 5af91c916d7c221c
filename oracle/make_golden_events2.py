#!/usr/bin/env python
"""tests/golden/events2_*.npz: the UNMODIFIED reference run with NON-AFFINE event functions (oracle/event_cases.py), one
OdeSystem per trajectory -- the counterpart of make_golden_events.py for the event-search kernels.  TEST INFRASTRUCTURE,
build container only.

    OMP_NUM_THREADS=1 python oracle/make_golden_events2.py

Per trajectory: the reference's event list (time, state, event index), final time / state, accepted steps, status, and
the number of sign changes of every g over the reference's own accepted steps (the reference reports a subset of
those: its root finder calls a root found only if |g(root)| <= 4 eps ABSOLUTE, utilities/optimizer.py:194-197)."""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

os.environ.setdefault("OMP_NUM_THREADS", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, "refshim"), "/root/reference", ROOT]

import numpy as np  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MAX_EV = 64


def run_one(job):
    import desolver as de
    from desolver_b200 import rhs_plugins
    from oracle import event_cases as ec
    case, y0, consts = job
    rhs_name, evs, tf, tol, method = ec.CASES[case]
    rhs = rhs_plugins.BUILTIN[rhs_name]
    a = de.OdeSystem(rhs, y0=np.array(y0, dtype=np.float64), t=(0.0, tf), dt=1e-2, rtol=tol, atol=tol, constants=dict(consts))
    a.method = method
    a.integrate(events=list(evs))
    dim = len(y0)
    rec = np.full((MAX_EV, dim + 2), np.nan)
    for k, st in enumerate(a.events[:MAX_EV]):
        rec[k, 0] = float(st.t)
        rec[k, 1:1 + dim] = np.asarray(st.y, dtype=np.float64)
        rec[k, -1] = evs.index(st.event)
    ys, ts = np.asarray(a.y), np.asarray(a.t)
    crossings = np.zeros(len(evs), dtype=np.int32)
    for e, ev in enumerate(evs):
        if ev.requires_dstate:
            gv = np.array([ev(ts[k], ys[k], rhs(ts[k], ys[k], **dict(consts)), **dict(consts)) for k in range(len(ts))])
        else:
            gv = np.array([ev(ts[k], ys[k], **dict(consts)) for k in range(len(ts))])
        chg = gv[:-1] * gv[1:] < 0
        if ev.direction > 0:
            chg &= gv[:-1] < 0
        elif ev.direction < 0:
            chg &= gv[:-1] > 0
        crossings[e] = int(chg.sum())
    return dict(rec=rec, count=len(a.events), y=np.asarray(a.y[-1], dtype=np.float64), t=float(a.t[-1]), accepted=len(a) - 1,
                status=2 if "terminated" in a.integration_status else 1, crossings=crossings)


def main():
    from desolver_b200 import benchmarks as bm
    from oracle import event_cases as ec
    L = bm.LORENZ
    lconst = dict(sigma=L["sigma"], rho=L["rho"], beta=L["beta"])
    n = 64
    yl = bm.lorenz_ensemble(n, seed=31)
    yv, mu = bm.vdp_ensemble(n, seed=32)
    with ProcessPoolExecutor(os.cpu_count()) as pool:
        for case, (rhs_name, evs, tf, tol, method) in ec.CASES.items():
            y0 = yl if rhs_name == "lorenz" else yv
            jobs = [(case, y0[:, i], lconst if rhs_name == "lorenz" else dict(mu=float(mu[i]))) for i in range(n)]
            res = list(pool.map(run_one, jobs, chunksize=2))
            blob = dict(y0=y0, records=np.stack([r["rec"] for r in res]), count=np.array([r["count"] for r in res], dtype=np.int32),
                        y_final=np.stack([r["y"] for r in res], axis=-1), t_final=np.array([r["t"] for r in res]),
                        accepted=np.array([r["accepted"] for r in res], dtype=np.int32),
                        status=np.array([r["status"] for r in res], dtype=np.int32),
                        crossings=np.stack([r["crossings"] for r in res]))
            if rhs_name != "lorenz":
                blob["mu"] = mu
            np.savez_compressed(os.path.join(GOLD, "events2_%s.npz" % case), **blob)
            print("%s: reference events per trajectory mean %.1f max %d; sign changes mean %.1f; terminated %d of %d" %
                  (case, blob["count"].mean(), blob["count"].max(), blob["crossings"].sum(1).mean(), int((blob["status"] == 2).sum()), n))
    # the reference's own multiple-roots test: one oscillator, a cubic in t, step-size cap through a callback
    import desolver as de
    from desolver_b200 import rhs_plugins
    a = de.OdeSystem(rhs_plugins.harmonic_oscillator, y0=np.array([1.0, 0.0]), t=(0, np.pi / 4), dt=np.pi / 512, rtol=1e-8, atol=1e-8,
                     constants=dict(k=1.0, m=1.0))
    a.method = "RK45"

    def dt_max_cb(ode_sys):
        ode_sys.dt = min(ode_sys.dt, np.pi / 64)
    a.integrate(events=ec.time_cubic, callback=dt_max_cb)
    np.savez_compressed(os.path.join(GOLD, "events2_time_cubic.npz"), t=np.array([float(e.t) for e in a.events]),
                        y=np.stack([np.asarray(e.y) for e in a.events]), steps=np.int64(len(a) - 1), t_final=float(a.t[-1]))
    print("time_cubic: events at", [float(e.t) for e in a.events], "steps", len(a) - 1)


if __name__ == "__main__":
    main()
