#!/usr/bin/env python
"""Round-2 fixtures from the UNMODIFIED reference (Microno95/desolver v5.1.0), same rules as make_golden.py:
TEST INFRASTRUCTURE, runs only in the build container (needs /root/reference and oracle/refshim).

    OMP_NUM_THREADS=1 python oracle/make_golden_round2.py [--only history,fixed,...]

  history_sho.npz        every accepted step (a.t, a.y) of the README oscillator for RK45 / RK87 / BABS9O7H, time
                         slices a[t0:t1], nearest-step lookup a[t], array-valued sol(t) and sol.grad(t)
  fixed_step.npz         Lorenz trajectories through the fixed-step tableaux that had no fixture (RK5, Midpoint,
                         Heun's, Ralston's, Euler, Euler-Trapezoidal) and the Symplectic Euler splitting scheme
  solve_ivp_teval.npz    solve_ivp(..., t_eval=...) of the reference (repeated integrate(t): exact stops)
  richardson_sho.npz     generate_richardson_integrator over RK4 / RK45CK / Midpoint on the README oscillator: every step
"""
import argparse
import os
import sys

os.environ.setdefault("OMP_NUM_THREADS", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, "refshim"), "/root/reference", ROOT]

import numpy as np  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import desolver as de
    from desolver_b200 import benchmarks as bm
    from desolver_b200 import rhs_plugins

    def want(name):
        return not args.only or name in args.only.split(",")

    if want("history"):
        blob = {}
        for key, method, kw in (("rk45", "RK45", dict(dt=0.01, rtol=1e-9, atol=1e-9)),
                                ("rk87", "RK87", dict(dt=0.01, rtol=1e-9, atol=1e-9)),
                                ("babs9o7h", "BABS9O7H", dict(dt=0.05))):
            a = de.OdeSystem(rhs_plugins.harmonic_oscillator, y0=np.array([1.0, 0.0]), t=(0, 2 * np.pi),
                             constants=dict(k=1.0, m=1.0), dense_output=(key == "rk45"), **kw)
            a.method = method
            a.integrate()
            blob[key + ".t"] = np.asarray(a.t)
            blob[key + ".y"] = np.asarray(a.y)
            blob[key + ".len"] = np.int64(len(a))
            if key == "rk45":
                s = a[0.5:2.0]
                blob["slice.t"], blob["slice.y"] = np.asarray(s.t), np.asarray(s.y)
                s = a[1.0:]
                blob["slice_open.t"] = np.asarray(s.t)
                s = a[:3.0:2]
                blob["slice_step.t"] = np.asarray(s.t)
                tq = np.array([0.0, 0.005, 0.3, 1.0, 3.0, 6.0, 2 * np.pi])
                blob["tq"] = tq
                blob["sol"] = np.asarray(a.sol(tq))                                  # (T, dim)
                blob["sol2d"] = np.asarray(a.sol(tq[:6].reshape(2, 3)))               # (2, 3, dim)
                blob["grad"] = np.stack([np.asarray(a.sol.grad(float(x))) for x in tq])
                # a second integrate() call continues the same history
                a.integrate(7.0)
                blob["cont.t"] = np.asarray(a.t)
                blob["cont.y"] = np.asarray(a.y)
        # nearest-step lookup without dense output (reference :1187-1196)
        a = de.OdeSystem(rhs_plugins.harmonic_oscillator, y0=np.array([1.0, 0.0]), t=(0, 2 * np.pi), dt=0.01, rtol=1e-9,
                         atol=1e-9, constants=dict(k=1.0, m=1.0))
        a.integrate()
        near = [1.0, 2.5, 6.2]
        blob["near.q"] = np.array(near)
        blob["near.t"] = np.array([float(a[x].t) for x in near])
        np.savez_compressed(os.path.join(GOLD, "history_sho.npz"), **blob)
        print("history: rk45 %d rows, rk87 %d, babs %d, continued %d" % (blob["rk45.len"], blob["rk87.len"],
                                                                         blob["babs9o7h.len"], len(blob["cont.t"])))

    if want("fixed"):
        L = bm.LORENZ
        lconst = dict(sigma=L["sigma"], rho=L["rho"], beta=L["beta"])
        n = 8
        y0 = bm.lorenz_ensemble(n, seed=4)
        blob = dict(y0=y0, tf=0.25, dt0=2e-3)
        for method in ("RK5", "Midpoint", "Heun's", "Ralston's", "Euler", "Euler-Trapezoidal"):
            ys, ts, steps, nfev = [], [], [], []
            for i in range(n):
                a = de.OdeSystem(rhs_plugins.lorenz, y0=y0[:, i].copy(), t=(0.0, 0.25), dt=2e-3, constants=lconst)
                a.method = method
                a.integrate()
                ys.append(np.asarray(a.y[-1]))
                ts.append(float(a.t[-1]))
                steps.append(len(a) - 1)
                nfev.append(int(a.nfev))
            blob[method + ".y"] = np.stack(ys, axis=-1)
            blob[method + ".t"] = np.array(ts)
            blob[method + ".steps"] = np.array(steps)
            blob[method + ".nfev"] = np.array(nfev)
        # Symplectic Euler on the oscillator (state (2,): q, p)
        a = de.OdeSystem(rhs_plugins.harmonic_oscillator, y0=np.array([1.0, 0.0]), t=(0, 1.0), dt=0.01, constants=dict(k=1.0, m=1.0))
        a.method = "Symplectic Forward Euler"
        a.integrate()
        blob["symp_euler.t"] = np.asarray(a.t)
        blob["symp_euler.y"] = np.asarray(a.y)
        np.savez_compressed(os.path.join(GOLD, "fixed_step.npz"), **blob)
        print("fixed-step fixtures:", {k: int(blob[k + ".steps"][0]) for k in ("RK5", "Euler")})

    if want("richardson"):
        # local Richardson extrapolation (integrator_types.py:420-613) over a fixed-step and an adaptive basis scheme
        blob = {}
        for key, base, it in (("rk4_2", de.integrators.RK4Solver, 2), ("rk4_3", de.integrators.RK4Solver, 3),
                              ("rk45ck_2", de.integrators.RK45CKSolver, 2), ("midpoint_4", de.integrators.MidpointSolver, 4)):
            R = de.integrators.generate_richardson_integrator(base, richardson_iter=it)
            a = de.OdeSystem(rhs_plugins.harmonic_oscillator, y0=np.array([1.0, 0.0]), t=(0, 2 * np.pi), dt=0.1, rtol=1e-9, atol=1e-9,
                             constants=dict(k=1.0, m=1.0))
            a.set_method(R)
            a.integrate()
            blob[key + ".t"], blob[key + ".y"], blob[key + ".nfev"] = np.asarray(a.t), np.asarray(a.y), np.int64(a.nfev)
            print("richardson %s: %d steps, nfev %d" % (key, len(a) - 1, a.nfev))
        np.savez_compressed(os.path.join(GOLD, "richardson_sho.npz"), **blob)

    if want("vector_tol"):
        # array-valued tolerances (integrator_types.py:37-38): one (rtol, atol) per state component
        L = bm.LORENZ
        lconst = dict(sigma=L["sigma"], rho=L["rho"], beta=L["beta"])
        y0 = bm.lorenz_ensemble(12, seed=14)
        rtol, atol = np.array([1e-9, 1e-7, 1e-8]), np.array([1e-9, 1e-10, 1e-6])
        ys, acc, att = [], [], []
        for i in range(y0.shape[1]):
            a = de.OdeSystem(rhs_plugins.lorenz, y0=y0[:, i].copy(), t=(0.0, 2.0), dt=1e-2, rtol=rtol, atol=atol, constants=lconst)
            a.method = "RK45"
            calls = [0]
            inner = a.integrator.step

            def counted(*p, _inner=inner, **k):
                calls[0] += 1
                return _inner(*p, **k)
            a.integrator.step = counted
            a.integrate()
            ys.append(np.asarray(a.y[-1]))
            acc.append(len(a) - 1)
            att.append(calls[0])
        np.savez_compressed(os.path.join(GOLD, "lorenz_vector_tol.npz"), y0=y0, rtol=rtol, atol=atol, y_final=np.stack(ys, axis=-1),
                            accepted=np.array(acc, dtype=np.int32), attempts=np.array(att, dtype=np.int32))
        print("vector_tol: accepted", acc)

    if want("teval"):
        L = bm.LORENZ
        y0 = bm.lorenz_ensemble(6, seed=12)
        t_eval = np.array([0.1, 0.25, 0.7, 1.0, 1.3])
        ys, ts, nfev = [], [], []
        for i in range(y0.shape[1]):
            r = de.solve_ivp(rhs_plugins.lorenz, (0.0, 1.5), y0[:, i].copy(), method="RK45", t_eval=t_eval, rtol=1e-9, atol=1e-9,
                             first_step=1e-2, args=(L["sigma"], L["rho"], L["beta"]))
            ys.append(np.asarray(r.y))
            ts.append(np.asarray(r.t))
            nfev.append(int(r.nfev))
        np.savez_compressed(os.path.join(GOLD, "solve_ivp_teval.npz"), y0=y0, t_eval=t_eval, y=np.stack(ys, axis=1),
                            t=np.stack(ts, axis=0), nfev=np.array(nfev))
        print("t_eval fixture: y", np.stack(ys, axis=1).shape)


if __name__ == "__main__":
    main()
