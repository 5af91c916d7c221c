#!/usr/bin/env python
"""Generate tests/golden/events_*.npz by running the UNMODIFIED reference (Microno95/desolver v5.1.0) with event
detection, one `OdeSystem` per trajectory (the reference has no ensemble concept).

TEST INFRASTRUCTURE, build container only (needs /root/reference; `autoray` is the dispatch-only stand-in of
oracle/refshim).  The event objects are desolver_b200.events.HyperplaneEvent instances: plain callables with the
reference's `is_terminal` / `direction` attributes, so the reference consumes them unchanged
(/root/reference/desolver/differential_system.py:29-57).

    OMP_NUM_THREADS=1 python oracle/make_golden_events.py [--procs 8]

Per trajectory the fixture holds the reference's event list (time, state, event index), the final time / state,
the accepted-step count and the integration status string class (1 done, 2 terminated by an event).

Note on the reference's detection rule: `brentsrootvec` reports success only if |g(root)| <= 4 eps ABSOLUTE
(utilities/optimizer.py:194-197, 316), so whether an event is reported depends on the last-ulp noise of the
interpolant at the root when |state| >> |g|; the fixtures therefore use events whose g is a bare state component
(well resolved near zero), and the parity test treats reference events as a SUBSET requirement."""
import argparse
import os
import sys
from concurrent.futures import ProcessPoolExecutor

os.environ.setdefault("OMP_NUM_THREADS", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, "refshim"), "/root/reference", ROOT]

import numpy as np  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MAX_EV = 48


def event_set(name):
    from desolver_b200.events import component_crossing, component_extremum
    if name == "lorenz_zmax":          # successive maxima of z (the Lorenz map): dz/dt falls through 0 -- requires_dstate
        return [component_extremum(2, dim=3, direction=-1), component_crossing(0, 0.0, dim=3, direction=0)]
    if name == "lorenz_sections":      # lobe switches (x = 0, both directions) and falling crossings of y = 0
        return [component_crossing(0, 0.0, dim=3, direction=0), component_crossing(1, 0.0, dim=3, direction=-1)]
    if name == "lorenz_terminal":      # stop at the first rising crossing of x = 0; y = 0 crossings are recorded until then
        return [component_crossing(0, 0.0, dim=3, direction=1, is_terminal=True), component_crossing(1, 0.0, dim=3, direction=0)]
    if name == "vdp_sections":         # x = 0 falling, v = 0 both
        return [component_crossing(0, 0.0, dim=2, direction=-1), component_crossing(1, 0.0, dim=2, direction=0)]
    if name == "vdp_terminal":         # stop at the second kind of event: v = 0 rising is terminal
        return [component_crossing(0, 0.0, dim=2, direction=0), component_crossing(1, 0.0, dim=2, direction=1, is_terminal=True)]
    raise KeyError(name)


def run_one(job):
    import desolver as de
    from desolver_b200 import rhs_plugins
    rhs_name, evname, y0, tf, dt0, tol, consts = job
    evs = event_set(evname)
    a = de.OdeSystem(rhs_plugins.BUILTIN[rhs_name], y0=np.array(y0, dtype=np.float64), t=(0.0, tf), dt=dt0, rtol=tol, atol=tol,
                     constants=dict(consts))
    a.method = "RK45"
    a.integrate(events=evs)
    dim = len(y0)
    rec = np.full((MAX_EV, dim + 2), np.nan)
    for k, st in enumerate(a.events[:MAX_EV]):
        rec[k, 0] = float(st.t)
        rec[k, 1:1 + dim] = np.asarray(st.y, dtype=np.float64)
        rec[k, -1] = evs.index(st.event)
    status = 2 if "terminated" in a.integration_status else 1
    # independent count from the reference's own step history: accepted steps over which g changes sign in an allowed
    # direction (the bracketing test of utilities/optimizer.py:252).  The reference reports a subset of these.
    ys, ts = np.asarray(a.y), np.asarray(a.t)
    crossings = np.zeros(len(evs), dtype=np.int32)
    for e, ev in enumerate(evs):
        if getattr(ev, "requires_dstate", False):
            gv = np.array([ev(ts[k], ys[k], rhs_plugins.BUILTIN[rhs_name](ts[k], ys[k], **dict(consts))) for k in range(len(ts))])
        else:
            gv = np.array([ev(ts[k], ys[k]) for k in range(len(ts))])
        chg = gv[:-1] * gv[1:] < 0
        if ev.direction > 0:
            chg &= gv[:-1] < 0
        elif ev.direction < 0:
            chg &= gv[:-1] > 0
        crossings[e] = int(chg.sum())
    return dict(rec=rec, count=len(a.events), y=np.asarray(a.y[-1], dtype=np.float64), t=float(a.t[-1]), accepted=len(a) - 1,
                status=status, crossings=crossings)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from desolver_b200 import benchmarks as bm
    L = bm.LORENZ
    lconst = dict(sigma=L["sigma"], rho=L["rho"], beta=L["beta"])
    n = 96
    yl = bm.lorenz_ensemble(n, seed=21)
    yv, mu = bm.vdp_ensemble(n, seed=22)
    cases = {
        "lorenz_zmax": [("lorenz", "lorenz_zmax", yl[:, i], 2.0, 1e-2, 1e-9, lconst) for i in range(n)],
        "lorenz_sections": [("lorenz", "lorenz_sections", yl[:, i], 2.0, 1e-2, 1e-9, lconst) for i in range(n)],
        "lorenz_terminal": [("lorenz", "lorenz_terminal", yl[:, i], 2.0, 1e-2, 1e-9, lconst) for i in range(n)],
        "vdp_sections": [("van_der_pol", "vdp_sections", yv[:, i], 10.0, 1e-2, 1e-9, dict(mu=float(mu[i]))) for i in range(n)],
        "vdp_terminal": [("van_der_pol", "vdp_terminal", yv[:, i], 10.0, 1e-2, 1e-9, dict(mu=float(mu[i]))) for i in range(n)],
    }
    with ProcessPoolExecutor(args.procs) as pool:
        for name, jobs in cases.items():
            if args.only and name not in args.only.split(","):
                continue
            res = list(pool.map(run_one, jobs, chunksize=4))
            y0 = yl if name.startswith("lorenz") else yv
            blob = dict(y0=y0, records=np.stack([r["rec"] for r in res]), count=np.array([r["count"] for r in res], dtype=np.int32),
                        y_final=np.stack([r["y"] for r in res], axis=-1), t_final=np.array([r["t"] for r in res]),
                        accepted=np.array([r["accepted"] for r in res], dtype=np.int32),
                        status=np.array([r["status"] for r in res], dtype=np.int32),
                        crossings=np.stack([r["crossings"] for r in res]))
            if not name.startswith("lorenz"):
                blob["mu"] = mu
            np.savez_compressed(os.path.join(GOLD, "events_%s.npz" % name), **blob)
            print("%s: events per trajectory mean %.1f max %d (sign changes in the step history: mean %.1f), terminated %d of %d" %
                  (name, blob["count"].mean(), blob["count"].max(), blob["crossings"].sum(1).mean(),
                   int((blob["status"] == 2).sum()), n))


if __name__ == "__main__":
    main()
