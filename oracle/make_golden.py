#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference (Microno95/desolver v5.1.0).

TEST INFRASTRUCTURE.  Runs only in the build container, where /root/reference exists; the
reference's missing `autoray` dependency is replaced by the dispatch-only stand-in in
oracle/refshim (no arithmetic of its own).  The GPU box never runs this: it consumes the
committed fixtures.

    OMP_NUM_THREADS=1 python oracle/make_golden.py [--procs 8]

Each ensemble fixture holds, per trajectory, what ONE `de.OdeSystem(rhs, y0[:, i], ...)` of the
reference produced: final state, final time, accepted steps, step attempts (calls of
`integrator.step`), the last dt.  Per-trajectory runs are the parity oracle for the ensemble
engine.
"""
import argparse
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

os.environ.setdefault("OMP_NUM_THREADS", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, "refshim"), "/root/reference", ROOT]

import numpy as np  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _ref():
    import desolver as de
    return de


def run_one(job):
    """job = (rhs_name, method, y0, t0, tf, dt0, rtol, atol, constants, dense_times)"""
    de = _ref()
    from desolver_b200 import rhs_plugins
    rhs_name, method, y0, t0, tf, dt0, rtol, atol, consts, dense_times = job
    rhs = rhs_plugins.BUILTIN[rhs_name]
    a = de.OdeSystem(rhs, y0=np.array(y0, dtype=np.float64), t=(t0, tf), dt=dt0, rtol=rtol, atol=atol,
                     constants=dict(consts), dense_output=dense_times is not None)
    a.method = method
    calls = [0]
    inner = a.integrator.step

    def counted(*p, **k):
        calls[0] += 1
        return inner(*p, **k)
    a.integrator.step = counted
    status = 0
    try:
        a.integrate()
    except de.exception_types.FailedIntegration:
        status = 1
    out = dict(y=np.asarray(a.y[-1], dtype=np.float64), t=float(a.t[-1]), accepted=len(a) - 1,
               attempts=calls[0], dt=float(a.dt), nfev=int(a.nfev), status=status)
    if dense_times is not None and status == 0:
        out["dense"] = np.stack([np.asarray(a.sol(float(tq))) for tq in dense_times])
        out["n_interp"] = len(a.sol.y_interpolants)
    return out


def run_many(pool, jobs):
    return list(pool.map(run_one, jobs, chunksize=max(1, len(jobs) // 64)))


def pack(results, extra):
    d = dict(extra)
    d["y_final"] = np.stack([r["y"] for r in results], axis=-1)
    d["t_final"] = np.array([r["t"] for r in results])
    d["dt_final"] = np.array([r["dt"] for r in results])
    d["accepted"] = np.array([r["accepted"] for r in results], dtype=np.int32)
    d["attempts"] = np.array([r["attempts"] for r in results], dtype=np.int32)
    d["nfev"] = np.array([r["nfev"] for r in results], dtype=np.int64)
    d["status"] = np.array([r["status"] for r in results], dtype=np.int32)
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    de = _ref()
    from desolver_b200 import benchmarks as bm
    from desolver_b200.integrators import tableaux as tb

    def want(name):
        return not args.only or name in args.only.split(",")

    # ---- tableaux: the reference class data, and proof that ours are bit-identical ----
    if want("tableaux"):
        blob = {}
        registry = {}
        for cls in de.integrators.explicit_methods():
            name = cls.__name__
            blob[name + ".inter"] = np.asarray(cls.tableau_intermediate, dtype=np.float64)
            blob[name + ".order"] = np.float64(cls.__order__)
            if hasattr(cls, "tableau_final") and cls.tableau_final is not None:
                blob[name + ".final"] = np.asarray(cls.tableau_final, dtype=np.float64)
                mine = tb.rk_arrays(name)
                assert mine[0].tobytes() == blob[name + ".inter"].tobytes(), name
                assert mine[1].tobytes() == blob[name + ".final"].tobytes(), name
            else:
                mine = tb.symplectic_array(name)
                assert mine[0].tobytes() == blob[name + ".inter"].tobytes(), name
            registry[name] = list(getattr(cls, "__alt_names__", ()))
        np.savez_compressed(os.path.join(GOLD, "tableaux.npz"), **blob)
        json.dump(registry, open(os.path.join(GOLD, "method_names.json"), "w"), indent=1, sort_keys=True)
        print("tableaux: %d classes bit-identical" % len(registry))

    with ProcessPoolExecutor(args.procs) as pool:
        L = bm.LORENZ
        lconst = dict(sigma=L["sigma"], rho=L["rho"], beta=L["beta"])
        if want("lorenz_rk45"):
            n = 1024
            y0 = bm.lorenz_ensemble(n)
            jobs = [("lorenz", "RK45", y0[:, i], L["t0"], L["tf"], L["dt0"], L["rtol"], L["atol"], lconst, None)
                    for i in range(n)]
            res = run_many(pool, jobs)
            np.savez_compressed(os.path.join(GOLD, "lorenz_rk45.npz"),
                                **pack(res, dict(y0=y0, tf=L["tf"], dt0=L["dt0"], rtol=L["rtol"], atol=L["atol"])))
            att = np.array([r["attempts"] for r in res])
            print("lorenz_rk45: attempts mean %.1f min %d max %d" % (att.mean(), att.min(), att.max()))
        if want("lorenz_rk87"):
            n = 256
            y0 = bm.lorenz_ensemble(n)
            jobs = [("lorenz", "RK87", y0[:, i], L["t0"], L["tf"], L["dt0"], L["rtol"], L["atol"], lconst, None)
                    for i in range(n)]
            res = run_many(pool, jobs)
            np.savez_compressed(os.path.join(GOLD, "lorenz_rk87.npz"),
                                **pack(res, dict(y0=y0, tf=L["tf"], dt0=L["dt0"], rtol=L["rtol"], atol=L["atol"])))
            print("lorenz_rk87 done")
        if want("lorenz_other"):
            # FSAL (DOPRI45), low-order adaptive (Heun-Euler), fixed step (RK4), high order (RK108)
            n = 32
            y0 = bm.lorenz_ensemble(n)
            for method, key, tol, tf in (("Dormand-Prince", "dopri45", 1e-9, 2.0), ("AHE", "heun_euler", 1e-4, 0.5),
                                         ("RK4", "rk4", 1e-9, 0.5), ("RK108", "rk108", 1e-11, 2.0),
                                         ("RK1412", "rk1412", 1e-12, 2.0)):
                jobs = [("lorenz", method, y0[:, i], 0.0, tf, 1e-2, tol, tol, lconst, None) for i in range(n)]
                res = run_many(pool, jobs)
                np.savez_compressed(os.path.join(GOLD, "lorenz_%s.npz" % key),
                                    **pack(res, dict(y0=y0, tf=tf, dt0=1e-2, rtol=tol, atol=tol, method=method)))
            print("lorenz_other done")
        if want("vdp_rk45"):
            V = bm.VDP
            n = 512
            y0, mu = bm.vdp_ensemble(n)
            jobs = [("van_der_pol", "RK45", y0[:, i], V["t0"], V["tf"], V["dt0"], V["rtol"], V["atol"],
                     dict(mu=float(mu[i])), None) for i in range(n)]
            res = run_many(pool, jobs)
            np.savez_compressed(os.path.join(GOLD, "vdp_rk45.npz"),
                                **pack(res, dict(y0=y0, mu=mu, tf=V["tf"], dt0=V["dt0"], rtol=V["rtol"], atol=V["atol"])))
            att = np.array([r["attempts"] for r in res])
            print("vdp_rk45: attempts mean %.1f min %d max %d" % (att.mean(), att.min(), att.max()))
        if want("duffing_rk45"):
            # Duffing oscillator of the reference's example notebooks, per-trajectory forcing amplitude and frequency
            rng = np.random.default_rng(11)
            n = 128
            y0 = np.ascontiguousarray(rng.uniform(-1.5, 1.5, size=(2, n)))
            gamma = rng.uniform(0.1, 0.5, size=n)
            omega = rng.uniform(0.8, 1.6, size=n)
            jobs = [("duffing", "RK45", y0[:, i], 0.0, 10.0, 1e-2, 1e-9, 1e-9,
                     dict(delta=0.2, alpha=-1.0, beta=1.0, gamma=float(gamma[i]), omega=float(omega[i])), None) for i in range(n)]
            res = run_many(pool, jobs)
            np.savez_compressed(os.path.join(GOLD, "duffing_rk45.npz"),
                                **pack(res, dict(y0=y0, gamma=gamma, omega=omega, tf=10.0, dt0=1e-2, rtol=1e-9, atol=1e-9)))
            att = np.array([r["attempts"] for r in res])
            print("duffing_rk45: attempts mean %.1f min %d max %d" % (att.mean(), att.min(), att.max()))
        if want("edge"):
            # oversized first steps (2 consecutive rejections -> 3-term controller), loose/tight
            # tolerances, dt0 > tf - t0 (the |tf-t|/2 clamp), very short spans
            rng = np.random.default_rng(7)
            n = 48
            y0 = np.ascontiguousarray(rng.uniform([-20, -25, 0], [20, 25, 50], size=(n, 3)).T)
            dt0 = np.array([[1.0, 0.5, 2.0, 5.0][i % 4] for i in range(n)])
            tol = np.array([[1e-12, 1e-9, 1e-6][i % 3] for i in range(n)])
            tf = np.array([[1.5, 0.3, 1e-3, 1e-9][(i // 12) % 4] for i in range(n)])
            jobs = [("lorenz", "RK45", y0[:, i], 0.0, float(tf[i]), float(dt0[i]), float(tol[i]), float(tol[i]),
                     lconst, None) for i in range(n)]
            res = run_many(pool, jobs)
            np.savez_compressed(os.path.join(GOLD, "edge_rk45.npz"),
                                **pack(res, dict(y0=y0, tf=tf, dt0=dt0, rtol=tol, atol=tol)))
            print("edge: max rejected in one run", max(r["attempts"] - r["accepted"] for r in res))
        if want("linear_rk87"):
            n, cols = 48, 16
            A, Y0 = bm.linear_system(n, cols)
            jobs = [("linear", "RK87", Y0[:, i], 0.0, 1.0, 1e-2, 1e-9, 1e-9, dict(A=A), None) for i in range(cols)]
            res = run_many(pool, jobs)
            np.savez_compressed(os.path.join(GOLD, "linear_rk87.npz"),
                                **pack(res, dict(A=A, y0=Y0, tf=1.0, dt0=1e-2, rtol=1e-9, atol=1e-9)))
            print("linear_rk87 done")
        if want("forced"):
            # time-dependent RHS, the reference's own test fixture (tests/common.py:27-72)
            rng = np.random.default_rng(21)
            n = 64
            y0 = np.ascontiguousarray(rng.uniform(-2, 2, size=(n, 2)).T)
            for method, key in (("RK45", "rk45"), ("RK87", "rk87")):
                jobs = [("forced_oscillator", method, y0[:, i], 0.0, 5.0, 1e-2, 1e-10, 1e-10, {}, None) for i in range(n)]
                res = run_many(pool, jobs)
                np.savez_compressed(os.path.join(GOLD, "forced_%s.npz" % key),
                                    **pack(res, dict(y0=y0, tf=5.0, dt0=1e-2, rtol=1e-10, atol=1e-10)))
            print("forced done")
        if want("dense"):
            tq = np.sort(np.random.default_rng(3).uniform(0.0, 2.0, size=24))
            y0 = bm.lorenz_ensemble(8)
            jobs = [("lorenz", "RK45", y0[:, i], 0.0, 2.0, 1e-2, 1e-9, 1e-9, lconst, tq) for i in range(8)]
            res = run_many(pool, jobs)
            d = pack(res, dict(y0=y0, tq=tq, tf=2.0, dt0=1e-2, rtol=1e-9, atol=1e-9))
            d["dense"] = np.stack([r["dense"] for r in res], axis=-1)        # (n_query, dim, n_traj)
            d["n_interp"] = np.array([r["n_interp"] for r in res])
            np.savez_compressed(os.path.join(GOLD, "lorenz_dense.npz"), **d)
            print("dense done")

    # ---- symplectic N-body: the batched reference call IS the per-system semantics ----
    if want("nbody"):
        from desolver_b200 import rhs_plugins
        state, m = bm.nbody_ensemble(6, bodies=8, seed=1)                  # (2, B, nb, 3)
        ens = np.ascontiguousarray(np.moveaxis(state, 1, -1))               # (2, nb, 3, B): ensemble axis last
        blob = dict(state0=state, m=m, eps2=1e-4, G=1.0, dt=0.01, tf=0.105)
        for method in ("BABS9O7H", "ABAS5O6H"):
            a = de.OdeSystem(rhs_plugins.nbody, y0=ens, t=(0.0, 0.105), dt=0.01, constants=dict(m=m, eps2=1e-4, G=1.0))
            a.method = method
            a.integrate()
            blob[method + ".y"] = np.ascontiguousarray(np.moveaxis(np.asarray(a.y[-1]), -1, 1))   # back to (2, B, nb, 3)
            blob[method + ".t"] = np.asarray(a.t)
            # per-system runs must coincide bit for bit with the batched run (fixed dt, no controller)
            for b in range(2):
                s = de.OdeSystem(rhs_plugins.nbody, y0=state[:, b], t=(0.0, 0.105), dt=0.01,
                                 constants=dict(m=m, eps2=1e-4, G=1.0))
                s.method = method
                s.integrate()
                assert np.array_equal(s.y[-1], blob[method + ".y"][:, b]), "batched != per-system"
        np.savez_compressed(os.path.join(GOLD, "nbody_symplectic.npz"), **blob)
        print("nbody done: steps", len(blob["BABS9O7H.t"]) - 1)

    # ---- known answers ----
    if want("known"):
        from desolver_b200 import rhs_plugins
        hx = lambda v: [float(x).hex() for x in np.ravel(v)]
        ka = {}

        def readme(t, state, k, m, **kwargs):
            return np.array([[0.0, 1.0], [-k / m, 0.0]]) @ state
        for key, method in (("KA0", "RK45"), ("KA3", "RK87")):
            a = de.OdeSystem(readme, y0=np.array([1.0, 0.0]), t=(0, 2 * np.pi), dt=0.01, rtol=1e-9, atol=1e-9,
                             constants=dict(k=1.0, m=1.0))
            a.method = method
            a.integrate()
            b = de.OdeSystem(rhs_plugins.harmonic_oscillator, y0=np.array([1.0, 0.0]), t=(0, 2 * np.pi), dt=0.01,
                             rtol=1e-9, atol=1e-9, constants=dict(k=1.0, m=1.0))
            b.method = method
            b.integrate()
            assert np.array_equal(a.y, b.y) and np.array_equal(a.t, b.t), "plugin SHO != README matrix form"
            ka[key] = dict(accepted=len(a) - 1, nfev=int(a.nfev), y=hx(a.y[-1]), t=hx(a.t[-1]),
                           first_dts=hx(np.diff(a.t)[:4]), last_dt=hx(np.diff(a.t)[-1:]))
        a = de.OdeSystem(rhs_plugins.lorenz, y0=np.array([1.0, 1.0, 20.0]), t=(0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9)
        a.integrate()
        ka["KA1"] = dict(accepted=len(a) - 1, nfev=int(a.nfev), y=hx(a.y[-1]), first_dts=hx(np.diff(a.t)[:3]))
        for key, mu in (("KA2a", 0.1), ("KA2b", 10.0)):
            a = de.OdeSystem(rhs_plugins.van_der_pol, y0=np.array([2.0, 0.0]), t=(0, 10.0), dt=1e-2, rtol=1e-9,
                             atol=1e-9, constants=dict(mu=mu))
            a.integrate()
            ka[key] = dict(accepted=len(a) - 1, nfev=int(a.nfev), y=hx(a.y[-1]))
        for key, method in (("KA4", "BABS9O7H"), ("KA5", "ABAS5O6H")):
            a = de.OdeSystem(rhs_plugins.harmonic_oscillator, y0=np.array([1.0, 0.0]), t=(0, 2 * np.pi), dt=0.05)
            a.method = method
            a.integrate()
            ka[key] = dict(steps=len(a) - 1, nfev=int(a.nfev), y=hx(a.y[-1]), t=hx(a.t[-1]), last_dt=hx(np.diff(a.t)[-1:]))
        a = de.OdeSystem(rhs_plugins.harmonic_oscillator, y0=np.array([1.0, 0.0]), dense_output=True, t=(0, 2 * np.pi),
                         dt=0.01, rtol=1e-9, atol=1e-9, constants=dict(k=1.0, m=1.0))
        a.integrate()
        ka["KA6"] = dict(n_interp=len(a.sol.y_interpolants), sol_at_1=hx(a.sol(1.0)),
                         sol_at=[[tq, hx(a.sol(tq))] for tq in (0.0, 0.005, 0.01, 3.0, 6.28, 2 * np.pi)])
        json.dump(ka, open(os.path.join(GOLD, "known_answers.json"), "w"), indent=1, sort_keys=True)
        print(json.dumps(ka["KA0"]), json.dumps(ka["KA1"]))


if __name__ == "__main__":
    main()
