#!/usr/bin/env python
"""Differential run: the UNMODIFIED reference (numpy backend, from baseline/_ref) and desolver_b200 (cuda:0) solve the same
randomly drawn problems side by side on the GPU box, case by case:

  adaptive / fixed / symplectic / backward   one system, every explicit method family, plugin or torch-callable right-hand
                                             side, both arithmetic flavours: steps, step times, states, nfev, and the
                                             attributes a caller reads afterwards (dt, success, integration_status, ...)
  dense                                      + dense-output values, derivatives, time slices, nearest-step lookups
  continue / switch                          integrate(t1); integrate(); reset(); method / tolerance / dt / tf changes
                                             between calls; per-component tolerances
  solve_ivp / richardson                     the scipy-style front end (t_eval, first_step, max_step, args); Richardson
                                             wrappers over fixed-step and adaptive basis schemes
  shapes                                     matrix-valued states, 2000-component systems with the reference's global
                                             norm, non-zero start times, step-clamping callbacks
  ensemble / ensemble2 / events_ens          ensembles (fused and stage-level kernels, two legs, dense queries, events)
                                             against ONE reference OdeSystem per trajectory
  events / events2                           recorded and terminal events of arbitrary callables, with their states
  nbody                                      softened N-body through the fused symplectic kernel / the stage-level kernels

A case that deviates in the fast arithmetic flavour is re-run in STRICT arithmetic (strict_rerun_*).

    python tools/differential_vs_reference.py [--seed 0] [--cases 840] [--only kind,kind] > profiles/round2_differential_vs_reference.json

TEST / MEASUREMENT INFRASTRUCTURE: needs baseline/_ref (baseline/install_reference.py) and a GPU.  `autoray`, which the
offline image lacks, is the dispatch-only stand-in oracle/refshim."""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _import_reference():
    try:
        import autoray  # noqa: F401
    except Exception:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    with contextlib.redirect_stdout(io.StringIO()):
        import desolver
    return desolver


# ---- problems: the same callable body serves numpy (reference) and torch (engine) through `xp.stack` ----
def _rhs_factory(name, xp):
    if name == "sho":
        return lambda t, y, k, m: xp.stack([y[1], -k / m * y[0]])
    if name == "lorenz":
        return lambda t, y, sigma, rho, beta: xp.stack([sigma * (y[1] - y[0]), y[0] * (rho - y[2]) - y[1], y[0] * y[1] - beta * y[2]])
    if name == "vdp":
        return lambda t, y, mu: xp.stack([y[1], mu * (1.0 - y[0] * y[0]) * y[1] - y[0]])
    if name == "forced":
        return lambda t, y: xp.stack([y[1], t - y[0]])
    raise KeyError(name)


def _draw(rng, name):
    if name == "sho":
        return rng.uniform(-1, 1, 2), dict(k=float(rng.uniform(0.5, 4.0)), m=float(rng.uniform(0.5, 2.0))), (0.0, float(rng.uniform(2.0, 6.5)))
    if name == "lorenz":
        return rng.uniform([-15, -20, 5], [15, 20, 40]), dict(sigma=10.0, rho=28.0, beta=8.0 / 3.0), (0.0, float(rng.uniform(0.3, 1.5)))
    if name == "vdp":
        return rng.uniform(-2, 2, 2), dict(mu=float(10.0 ** rng.uniform(-1, 0.7))), (0.0, float(rng.uniform(1.0, 6.0)))
    if name == "forced":
        return rng.uniform(-1, 1, 2), dict(), (0.0, float(rng.uniform(1.0, 4.0)))
    raise KeyError(name)


ADAPTIVE = ["RK45", "RK87", "DOPRI45", "AHE", "RK108", "RK1412"]
FIXED = ["RK4", "RK5", "Midpoint", "Heun's", "Ralston's", "Euler", "Euler-Trap"]
SYMPLECTIC = ["BABS9O7H", "ABAS5O6H", "Symplectic Forward Euler"]
PLUGIN = {"sho": "harmonic_oscillator", "lorenz": "lorenz", "vdp": "van_der_pol", "forced": "forced"}


def run_case(ref, de, torch, rng, kind):
    prob = str(rng.choice(["sho", "lorenz", "vdp", "forced"]))
    y0, consts, span = _draw(rng, prob)
    method = str(rng.choice(ADAPTIVE if kind != "fixed" else FIXED))
    if kind == "symplectic":
        prob, method = "sho", str(rng.choice(SYMPLECTIC))
        y0, consts, span = _draw(rng, prob)
    backward = kind == "backward"
    if backward:
        span = (span[1], span[0])
    tol = float(10.0 ** rng.choice([-6, -8, -9, -11])) if method in ADAPTIVE else None
    if method == "AHE":                                          # order 2: keep the step count bounded
        tol = float(10.0 ** rng.choice([-4, -5, -6]))
    dt0 = float(rng.choice([1e-3, 1e-2, 0.05, 0.5])) if method in ADAPTIVE else float(rng.choice([0.01, 0.02, 0.05]))
    if backward:
        dt0 = -min(abs(dt0), 1e-3)       # the reference's retry chain is broken for dt < 0: avoid rejections
    path = str(rng.choice(["plugin", "callable"]))
    dense = kind == "dense"
    strict = bool(rng.integers(0, 2))
    desc = dict(kind=kind, problem=prob, method=method, tol=tol, dt0=dt0, span=span, path=path, strict=strict,
                y0=[float(v) for v in y0], constants=consts)
    kw = dict(t=span, dt=dt0, constants=consts, dense_output=dense)
    if tol is not None:
        kw.update(rtol=tol, atol=tol)
    # --- reference ---
    r = ref.OdeSystem(_rhs_factory(prob, np), y0=y0.copy(), **kw)
    r.method = method
    t0 = time.perf_counter()
    try:
        r.integrate()
    except Exception as e:                                       # noqa: BLE001
        if backward:                                             # fact 10: min(new, h0) with negative steps
            return dict(desc, skipped="reference failed on a backward retry: " + repr(e)[:80], steps_equal=True, final_rel_err=0.0)
        desc["ref_error"] = repr(e)[:120] + " | cause: " + repr(e.__cause__)[:160]
        r = None
    desc["ref_s"] = time.perf_counter() - t0
    # --- engine ---
    rhs = getattr(de.rhs_plugins, PLUGIN[prob]) if path == "plugin" else _rhs_factory(prob, torch)
    a = de.OdeSystem(rhs, torch.as_tensor(y0.copy()), device="cuda:0", strict=strict, **kw)
    a.method = method
    t0 = time.perf_counter()
    try:
        a.integrate()
    except Exception as e:                                       # noqa: BLE001
        desc["gpu_error"] = repr(e)[:120] + " | cause: " + repr(e.__cause__)[:160]
    torch.cuda.synchronize()
    desc["gpu_s"] = time.perf_counter() - t0
    if r is None or "gpu_error" in desc:
        # both sides must fail alike (the reference raises FailedIntegration when 64 retries do not meet the tolerances)
        both = r is None and "gpu_error" in desc
        return dict(desc, both_failed=both, steps_equal=both, final_rel_err=0.0 if both else float("inf"))
    t_ref, y_ref = np.asarray(r.t, dtype=np.float64), np.asarray(r.y, dtype=np.float64)
    t_gpu, y_gpu = a.t.cpu().numpy(), a.y.cpu().numpy()
    out = dict(desc, steps_ref=len(r) - 1, steps_gpu=len(a) - 1, nfev_ref=int(r.nfev), nfev_gpu=int(a.nfev))
    out["steps_equal"] = out["steps_ref"] == out["steps_gpu"]
    scale = max(1.0, float(np.max(np.abs(y_ref))))
    out["final_rel_err"] = float(np.max(np.abs(y_gpu[-1] - y_ref[-1])) / scale)
    out["final_t_err"] = float(abs(t_gpu[-1] - t_ref[-1]))
    if out["steps_equal"]:
        out["max_t_err"] = float(np.max(np.abs(t_gpu - t_ref)))
        out["max_y_rel_err"] = float(np.max(np.abs(y_gpu - y_ref)) / scale)
    # the attributes a caller reads afterwards
    api = {}
    for name in ("dt", "t0", "tf", "success", "integration_status", "counter", "njev", "dim", "rtol", "atol", "staggered_mask"):
        try:
            u, v = getattr(r, name), getattr(a, name)
            if name == "staggered_mask":
                api[name] = (u is None) == (v is None) or method in SYMPLECTIC
            elif u is None or v is None:
                api[name] = (u is None and v is None) or "ref %r gpu %r" % (u, v)
            elif name == "dt":
                ok = abs(float(u) - float(v)) <= 1e-6 * abs(float(u)) + 1e-300 if out["steps_equal"] else True
                api[name] = ok or "ref %r gpu %r" % (float(u), float(v))
            elif name in ("t0", "tf", "rtol", "atol"):
                api[name] = float(u) == float(v)
            elif name == "dim":
                api[name] = tuple(u) == tuple(v)
            else:
                api[name] = u == v
        except Exception as e:                                   # noqa: BLE001
            api[name] = "error: " + repr(e)[:80]
    out["api_mismatch"] = {k: v for k, v in api.items() if v is not True}
    if not strict and (not out["steps_equal"] or out["final_rel_err"] > max(1e-9, 1e-2 * (tol or 0.0))):
        # is it the arithmetic flavour?  the same case in STRICT arithmetic (the reference's rounding sequence)
        b = de.OdeSystem(rhs, torch.as_tensor(y0.copy()), device="cuda:0", strict=True, **kw)
        b.method = method
        b.integrate()
        out["strict_rerun_steps_equal"] = len(b) == len(r)
        out["strict_rerun_final_rel_err"] = float(np.max(np.abs(b.y[-1].cpu().numpy() - y_ref[-1])) / scale)
        if out["strict_rerun_steps_equal"]:
            out["strict_rerun_max_t_err"] = float(np.max(np.abs(b.t.cpu().numpy() - t_ref)))
    if dense:
        q = np.sort(rng.uniform(min(span), max(span), 9))
        s_ref = np.stack([np.asarray(r.sol(float(x))) for x in q])
        s_gpu = np.stack([a.sol(float(x)).cpu().numpy().reshape(-1) for x in q])
        out["dense_rel_err"] = float(np.max(np.abs(s_gpu - s_ref)) / scale)
        g_ref = np.stack([np.asarray(r.sol.grad(float(x))) for x in q])
        g_gpu = np.stack([a.sol.grad(float(x)).cpu().numpy().reshape(-1) for x in q])
        out["dense_grad_rel_err"] = float(np.max(np.abs(g_gpu - g_ref)) / max(1.0, float(np.max(np.abs(g_ref)))))
        if out["steps_equal"]:
            lo, hi = sorted((float(q[2]), float(q[6])))
            out["slice_len_equal"] = len(r[lo:hi].t) == len(a[lo:hi].t)
            out["nearest_t_err"] = float(abs(float(r[float(q[4])].t) - float(a[float(q[4])].t)))
    return out


def run_continue_case(ref, de, torch, rng):
    """integrate(t1); integrate(t2); reset(); integrate(): history and state of a continued, then repeated, run."""
    prob = str(rng.choice(["sho", "lorenz", "vdp"]))
    y0, consts, span = _draw(rng, prob)
    method = str(rng.choice(["RK45", "RK87", "DOPRI45", "RK4"]))
    tol = float(10.0 ** rng.choice([-7, -9]))
    path = str(rng.choice(["plugin", "callable"]))
    kw = dict(t=span, dt=0.01, constants=consts)
    if method != "RK4":
        kw.update(rtol=tol, atol=tol)
    t1 = span[0] + 0.37 * (span[1] - span[0])
    r = ref.OdeSystem(_rhs_factory(prob, np), y0=y0.copy(), **kw)
    r.method = method
    rhs = getattr(de.rhs_plugins, PLUGIN[prob]) if path == "plugin" else _rhs_factory(prob, torch)
    a = de.OdeSystem(rhs, torch.as_tensor(y0.copy()), device="cuda:0", **kw)
    a.method = method
    out = dict(kind="continue", problem=prob, method=method, tol=tol, path=path)
    lens, err = [], 0.0
    for step in ("t1", "tf", "reset"):
        for o in (r, a):
            if step == "t1":
                o.integrate(t1)
            elif step == "tf":
                o.integrate()
            else:
                o.reset()
                o.integrate()
        lens.append((len(r), len(a)))
        yr = np.asarray(r.y[-1], dtype=np.float64)
        err = max(err, float(np.max(np.abs(a.y[-1].cpu().numpy() - yr)) / max(1.0, float(np.max(np.abs(yr))))))
        err = max(err, abs(float(a.t[-1]) - float(r.t[-1])))
    out.update(lens=lens, steps_equal=all(x == y for x, y in lens), final_rel_err=err, nfev_ref=int(r.nfev), nfev_gpu=int(a.nfev))
    return out


def run_switch_case(ref, de, torch, rng):
    """Attribute setters between calls: method switch, new tolerances, new step size, new final time; array tolerances."""
    prob = str(rng.choice(["lorenz", "vdp", "sho"]))
    y0, consts, span = _draw(rng, prob)
    path = str(rng.choice(["plugin", "callable"]))
    m1, m2 = (str(v) for v in rng.choice(["RK45", "RK87", "DOPRI45", "RK4", "RK108"], 2, replace=False))
    tol1, tol2 = float(10.0 ** rng.choice([-6, -8])), float(10.0 ** rng.choice([-9, -10]))
    vec_tol = bool(rng.integers(0, 2))
    kw = dict(t=span, dt=0.01, constants=consts)
    dim = y0.shape[0]
    rt1 = tol1 * (1.0 + np.arange(dim)) if vec_tol else tol1
    t1 = span[0] + 0.4 * (span[1] - span[0])
    new_tf = span[1] + 0.3
    rhs = getattr(de.rhs_plugins, PLUGIN[prob]) if path == "plugin" else _rhs_factory(prob, torch)
    r = ref.OdeSystem(_rhs_factory(prob, np), y0=y0.copy(), rtol=rt1, atol=tol1, **kw)
    a = de.OdeSystem(rhs, torch.as_tensor(y0.copy()), device="cuda:0", rtol=torch.as_tensor(rt1) if vec_tol else rt1, atol=tol1, **kw)
    lens = []
    desc = dict(kind="switch", problem=prob, path=path, methods=[m1, m2], tols=[tol1, tol2], vector_rtol=vec_tol, tol=tol1,
                y0=[float(v) for v in y0], constants=consts, span=span)
    for o in (r, a):
        o.method = m1
        o.integrate(t1)
    lens.append((len(r), len(a)))
    for name, o in (("ref", r), ("gpu", a)):
        try:
            o.method = m2
            o.rtol = tol2
            o.atol = tol2
            o.dt = 0.005
            o.tf = new_tf
            o.integrate()
        except Exception as e:                                   # noqa: BLE001
            desc[name + "_error"] = repr(e)[:100] + " | cause: " + repr(e.__cause__)[:200]
    if "ref_error" in desc or "gpu_error" in desc:
        both = "ref_error" in desc and "gpu_error" in desc
        return dict(desc, lens=lens, both_failed=both, steps_equal=both, final_rel_err=0.0 if both else float("inf"),
                    t_ref=float(r.t[-1]), t_gpu=float(a.t[-1]), len_ref=len(r), len_gpu=len(a))
    lens.append((len(r), len(a)))
    yr = np.asarray(r.y[-1], dtype=np.float64)
    return dict(kind="switch", problem=prob, path=path, methods=[m1, m2], tols=[tol1, tol2], vector_rtol=vec_tol, lens=lens, tol=tol1,
                steps_equal=all(x == y for x, y in lens), nfev_ref=int(r.nfev), nfev_gpu=int(a.nfev),
                final_t_err=abs(float(a.t[-1]) - float(r.t[-1])),
                final_rel_err=float(np.max(np.abs(a.y[-1].cpu().numpy() - yr)) / max(1.0, float(np.max(np.abs(yr))))))


def run_solve_ivp_case(ref, de, torch, rng):
    """The scipy-style front end: t_eval stops, first_step / max_step options, args tuple."""
    prob = str(rng.choice(["lorenz", "vdp", "sho"]))
    y0, consts, span = _draw(rng, prob)
    method = str(rng.choice(["RK45", "RK87", "DOPRI45"]))
    tol = float(10.0 ** rng.choice([-6, -8, -10]))
    opts = dict(rtol=tol, atol=tol, first_step=float(rng.choice([1e-3, 1e-2, 0.1])))
    if rng.integers(0, 2):
        opts["max_step"] = float(rng.choice([0.05, 0.2]))
    t_eval = None
    if rng.integers(0, 3) > 0:
        t_eval = np.sort(rng.uniform(span[0], span[1], 6))
        t_eval[-1] = span[1]
    args = tuple(consts.values())
    path = str(rng.choice(["plugin", "callable"]))
    rr = ref.solve_ivp(_rhs_factory(prob, np), span, y0.copy(), method=method, t_eval=t_eval, args=args or None, **opts)
    rhs = getattr(de.rhs_plugins, PLUGIN[prob]) if path == "plugin" else _rhs_factory(prob, torch)
    kw = dict(opts)
    ga = de.solve_ivp(rhs, span, torch.as_tensor(y0.copy()).cuda(), method=method,
                      t_eval=None if t_eval is None else torch.as_tensor(t_eval).cuda(), args=args or None, **kw)
    t_r, y_r = np.asarray(rr.t, dtype=np.float64), np.asarray(rr.y, dtype=np.float64)
    t_g, y_g = ga.t.cpu().numpy(), ga.y.cpu().numpy()
    out = dict(kind="solve_ivp", problem=prob, method=method, tol=tol, path=path, t_eval=t_eval is not None, max_step=opts.get("max_step"),
               shape_ref=list(y_r.shape), shape_gpu=list(y_g.shape), nfev_ref=int(rr.nfev), nfev_gpu=int(ga.nfev))
    out["steps_equal"] = t_r.shape == t_g.shape and y_r.shape == y_g.shape
    if out["steps_equal"]:
        out["max_t_err"] = float(np.max(np.abs(t_r - t_g)))
        out["path_rel_err"] = float(np.max(np.abs(y_r - y_g)) / max(1.0, float(np.max(np.abs(y_r)))))
        out["final_rel_err"] = float(np.max(np.abs(y_r[..., -1] - y_g[..., -1])) / max(1.0, float(np.max(np.abs(y_r)))))
    else:
        out["final_rel_err"] = float("inf")
    return out


def run_richardson_case(ref, de, torch, rng):
    """Local Richardson extrapolation over a basis scheme (integrator_types.py:420-613)."""
    base = str(rng.choice(["RK4Solver", "MidpointSolver", "RK45CKSolver"]))
    it = int(rng.choice([2, 3]))
    y0, consts, span = _draw(rng, "sho")
    tol = float(10.0 ** rng.choice([-6, -8]))
    kw = dict(t=span, dt=0.05, rtol=tol, atol=tol, constants=consts)
    r = ref.OdeSystem(_rhs_factory("sho", np), y0=y0.copy(), **kw)
    r.method = ref.integrators.generate_richardson_integrator(getattr(ref.integrators, base), richardson_iter=it)
    r.integrate()
    path = str(rng.choice(["plugin", "callable"]))
    rhs = de.rhs_plugins.harmonic_oscillator if path == "plugin" else _rhs_factory("sho", torch)
    a = de.OdeSystem(rhs, torch.as_tensor(y0.copy()), device="cuda:0", **kw)
    a.method = de.integrators.generate_richardson_integrator(getattr(de.integrators, base), richardson_iter=it)
    a.integrate()
    yr = np.asarray(r.y[-1], dtype=np.float64)
    return dict(kind="richardson", basis=base, richardson_iter=it, tol=tol, path=path, steps_ref=len(r) - 1, steps_gpu=len(a) - 1,
                steps_equal=len(r) == len(a), nfev_ref=int(r.nfev), nfev_gpu=int(a.nfev),
                final_rel_err=float(np.max(np.abs(a.y[-1].cpu().numpy() - yr)) / max(1.0, float(np.max(np.abs(yr))))))


def run_ensemble_case(ref, de, torch, rng):
    """A small ensemble through the fused kernel against one reference OdeSystem per trajectory."""
    prob = str(rng.choice(["lorenz", "vdp"]))
    method = str(rng.choice(["RK45", "RK87", "DOPRI45"]))
    n, tol = 24, float(10.0 ** rng.choice([-7, -9]))
    draws = [_draw(rng, prob) for _ in range(n)]
    span = draws[0][2]
    y0 = np.stack([d[0] for d in draws], axis=1)
    consts = dict(draws[0][1])
    per = None
    if prob == "vdp":
        per = np.array([d[1]["mu"] for d in draws])
        consts = dict(mu=torch.as_tensor(per, device="cuda:0"))
    a = de.OdeSystem(getattr(de.rhs_plugins, PLUGIN[prob]), torch.as_tensor(y0), t=span, dt=1e-2, rtol=tol, atol=tol, constants=consts,
                     ensemble=True, device="cuda:0")
    a.method = method
    a.integrate()
    acc, rej, y = a.accepted.cpu().numpy(), a.rejected.cpu().numpy(), a[-1].y.cpu().numpy()
    same, err = 0, 0.0
    for i in range(n):
        c = dict(draws[0][1]) if per is None else dict(mu=float(per[i]))
        r = ref.OdeSystem(_rhs_factory(prob, np), y0=y0[:, i].copy(), t=span, dt=1e-2, rtol=tol, atol=tol, constants=c)
        r.method = method
        calls = [0]
        inner = r.integrator.step

        def counted(*p, _inner=inner, **k):
            calls[0] += 1
            return _inner(*p, **k)
        r.integrator.step = counted
        r.integrate()
        same += int(len(r) - 1 == acc[i] and calls[0] == acc[i] + rej[i])
        yr = np.asarray(r.y[-1], dtype=np.float64)
        err = max(err, float(np.max(np.abs(y[:, i] - yr)) / max(1.0, np.max(np.abs(yr)))))
    return dict(kind="ensemble", problem=prob, method=method, tol=tol, trajectories=n, counts_identical=same, final_rel_err=err,
                steps_equal=same == n)


def run_ensemble2_case(ref, de, torch, rng):
    """Ensembles beyond the plain fused run: torch-callable right-hand sides (stage-level kernels, t of shape (N,)),
    two-leg integrations, dense-output queries, a method switch -- against one reference OdeSystem per trajectory."""
    prob = str(rng.choice(["lorenz", "vdp", "forced"]))
    m1 = str(rng.choice(["RK45", "RK87", "DOPRI45", "RK4"]))
    m2 = str(rng.choice(["RK45", "RK87", "RK108"])) if rng.integers(0, 2) else m1
    path = str(rng.choice(["plugin", "callable"]))
    dense = bool(rng.integers(0, 2))
    n, tol = 12, float(10.0 ** rng.choice([-7, -9]))
    draws = [_draw(rng, prob) for _ in range(n)]
    span = draws[0][2]
    t1 = span[0] + 0.45 * (span[1] - span[0])
    y0 = np.stack([d[0] for d in draws], axis=1)
    consts, per = dict(draws[0][1]), None
    if prob == "vdp":
        per = np.array([d[1]["mu"] for d in draws])
        consts = dict(mu=torch.as_tensor(per, device="cuda:0"))
    rhs = getattr(de.rhs_plugins, PLUGIN[prob]) if path == "plugin" else _rhs_factory(prob, torch)
    kw = dict(t=span, dt=1e-2)
    tk = dict(rtol=tol, atol=tol)
    a = de.OdeSystem(rhs, torch.as_tensor(y0), constants=consts, ensemble=True, device="cuda:0", dense_output=dense, **kw, **tk)
    a.method = m1
    a.integrate(t1)
    a.method = m2
    a.integrate()
    acc, rej, y = a.accepted.cpu().numpy(), a.rejected.cpu().numpy(), a[-1].y.cpu().numpy()
    q = float(span[0] + 0.77 * (span[1] - span[0]))
    sq = a.sol(q).cpu().numpy() if dense else None
    same, err, derr = 0, 0.0, 0.0
    for i in range(n):
        c = dict(draws[0][1]) if per is None else dict(mu=float(per[i]))
        r = ref.OdeSystem(_rhs_factory(prob, np), y0=y0[:, i].copy(), constants=c, dense_output=dense, **kw, **tk)
        r.method = m1
        r.integrate(t1)
        r.method = m2
        r.integrate()
        same += int(len(r) - 1 == acc[i])
        yr = np.asarray(r.y[-1], dtype=np.float64)
        sc = max(1.0, float(np.max(np.abs(yr))))
        err = max(err, float(np.max(np.abs(y[:, i] - yr)) / sc))
        if dense:
            derr = max(derr, float(np.max(np.abs(sq[:, i] - np.asarray(r.sol(q)))) / sc))
    out = dict(kind="ensemble2", problem=prob, methods=[m1, m2], path=path, dense=dense, tol=tol, trajectories=n, counts_identical=same,
               final_rel_err=err, steps_equal=same == n)
    if dense:
        out["dense_rel_err"] = derr
    return out


def run_shape_case(ref, de, torch, rng):
    """State shapes beyond a short vector: matrix-valued states (Y' = A Y + Y B), a long single system with the reference's
    global error norm (a ring of coupled oscillators), non-zero start times, callbacks that count and clamp the step."""
    variant = str(rng.choice(["matrix", "ring", "callback"]))
    method = str(rng.choice(["RK45", "RK87", "DOPRI45", "RK4", "AHE"]))
    tol = float(10.0 ** rng.choice([-5, -7, -9])) if method != "AHE" else 1e-5
    t0 = float(rng.choice([0.0, -1.5, 3.0]))
    span = (t0, t0 + float(rng.uniform(0.5, 2.0)))
    kw = dict(t=span, dt=0.01)
    if method != "RK4":
        kw.update(rtol=tol, atol=tol)
    if variant == "matrix":
        A = rng.normal(size=(3, 3)) * 0.7 - 0.5 * np.eye(3)
        B = rng.normal(size=(4, 4)) * 0.5
        y0 = rng.normal(size=(3, 4))
        At, Bt = torch.as_tensor(A).cuda(), torch.as_tensor(B).cuda()
        f_np = lambda t, y: A @ y + y @ B                          # noqa: E731
        f_t = lambda t, y: At @ y + y @ Bt                         # noqa: E731
    elif variant == "ring":
        m = int(rng.choice([64, 257, 1000]))
        y0 = rng.normal(size=(2 * m,))
        kk = float(rng.uniform(0.5, 3.0))
        f_np = lambda t, y: np.concatenate([y[m:], kk * (np.roll(y[:m], 1) - 2 * y[:m] + np.roll(y[:m], -1)) - np.sin(t) * y[:m]])      # noqa: E731
        f_t = lambda t, y: torch.cat([y[m:], kk * (torch.roll(y[:m], 1) - 2 * y[:m] + torch.roll(y[:m], -1)) - torch.sin(t) * y[:m]])  # noqa: E731
    else:
        y0 = rng.uniform(-1, 1, 2)
        f_np = _rhs_factory("forced", np)
        f_t = _rhs_factory("forced", torch)
    r = ref.OdeSystem(f_np, y0=y0.copy(), **kw)
    a = de.OdeSystem(f_t, torch.as_tensor(y0.copy()), device="cuda:0", **kw)
    calls = {"ref": 0, "gpu": 0}
    cap = float(rng.choice([0.03, 0.1]))
    for name, o in (("ref", r), ("gpu", a)):
        o.method = method
        if variant == "callback":
            def cb(sys_, _name=name):
                calls[_name] += 1
                if abs(float(sys_.dt)) > cap:
                    sys_.dt = cap
            o.integrate(callback=cb)
        else:
            o.integrate()
    yr = np.asarray(r.y[-1], dtype=np.float64)
    out = dict(kind="shapes", variant=variant, method=method, tol=tol if method != "RK4" else None, span=span, shape=list(y0.shape),
               steps_ref=len(r) - 1, steps_gpu=len(a) - 1, steps_equal=len(r) == len(a), nfev_ref=int(r.nfev), nfev_gpu=int(a.nfev),
               final_t_err=abs(float(a.t[-1]) - float(r.t[-1])), shape_equal=tuple(a.y.shape) == tuple(np.asarray(r.y).shape),
               final_rel_err=float(np.max(np.abs(a.y[-1].cpu().numpy() - yr)) / max(1.0, float(np.max(np.abs(yr))))))
    if variant == "callback":
        out["callback_calls"] = [calls["ref"], calls["gpu"]]
        out["steps_equal"] = out["steps_equal"] and calls["ref"] == calls["gpu"]
    if not out["shape_equal"]:
        out["steps_equal"] = False
    return out


def run_event2_case(ref, de, torch, rng):
    """Event records in detail: times AND states of every recorded event, direction filters, a `requires_dstate` event
    (a maximum of a component), on a single Van der Pol system."""
    y0, consts, _ = _draw(rng, "vdp")
    level = float(rng.uniform(-1.0, 1.0))
    direction = int(rng.choice([-1, 0, 1]))

    def make(xp):
        def cross(t, y, **k):
            return y[0] - level
        cross.direction = direction

        def vmax(t, y, dy, **k):                 # dy = the interpolant's derivative: zero of the acceleration-free velocity
            return dy[0]
        vmax.requires_dstate = True
        vmax.direction = -1

        def product(t, y, **k):
            return y[0] * y[1] - 0.3
        return [cross, vmax, product]
    kw = dict(t=(0.0, 8.0), dt=1e-2, rtol=1e-9, atol=1e-9, constants=consts, dense_output=True)
    r = ref.OdeSystem(_rhs_factory("vdp", np), y0=y0.copy(), **kw)
    r.method = "RK45"
    r.integrate(events=make(np))
    a = de.OdeSystem(_rhs_factory("vdp", torch) if rng.integers(0, 2) else de.rhs_plugins.van_der_pol, torch.as_tensor(y0.copy()),
                     device="cuda:0", **kw)
    a.method = "RK45"
    a.integrate(events=make(torch))
    ev_r = sorted((float(e.t), np.asarray(e.y, dtype=np.float64)) for e in r.events)
    ev_g = sorted((float(e.t), e.y.cpu().numpy().reshape(-1)) for e in a.events)
    miss, yerr, terr = 0, 0.0, 0.0
    for t, y in ev_r:                                    # the reference reports a subset of the crossings
        j = int(np.argmin([abs(t - u) for u, _ in ev_g])) if ev_g else -1
        if j < 0 or abs(ev_g[j][0] - t) > 1e-8:
            miss += 1
            continue
        terr = max(terr, abs(ev_g[j][0] - t))
        yerr = max(yerr, float(np.max(np.abs(ev_g[j][1] - y))))
    return dict(kind="events2", direction=direction, events_ref=len(ev_r), events_gpu=len(ev_g), ref_events_missing_on_gpu=miss,
                event_t_err=terr, event_y_err=yerr, steps_ref=len(r) - 1, steps_gpu=len(a) - 1,
                steps_equal=len(r) == len(a) and miss == 0 and len(ev_g) >= len(ev_r),
                final_rel_err=max(yerr, float(np.max(np.abs(a.y[-1].cpu().numpy() - np.asarray(r.y[-1]))))))


def run_nbody_case(ref, de, torch, rng):
    """Softened gravitational N-body, state (2, nb, 3): the engine's `nbody` plugin (a dual numpy / torch callable) through
    the reference, against the engine's fused symplectic kernel / stage-level kernels."""
    nb = int(rng.choice([3, 8, 16]))
    method = str(rng.choice(SYMPLECTIC + ["RK45", "RK87"]))
    state = np.stack([rng.normal(size=(nb, 3)), 0.3 * rng.normal(size=(nb, 3))])
    m = rng.uniform(0.5, 1.5, nb) / nb
    eps2 = float(rng.choice([1e-2, 1e-4]))
    dt = float(rng.choice([0.01, 0.02]))
    tf = float(rng.uniform(0.2, 0.6))
    kw = dict(t=(0.0, tf), dt=dt)
    tol = None
    if method in ("RK45", "RK87"):
        tol = float(10.0 ** rng.choice([-7, -9]))
        kw.update(rtol=tol, atol=tol)
    r = ref.OdeSystem(de.rhs_plugins.nbody, y0=state.copy(), constants=dict(m=m, eps2=eps2, G=1.0), **kw)
    r.method = method
    r.integrate()
    fused = bool(rng.integers(0, 2))
    a = de.OdeSystem(de.rhs_plugins.nbody, torch.as_tensor(state.copy()), device="cuda:0",
                     constants=dict(m=torch.as_tensor(m), eps2=eps2, G=1.0), strict=bool(rng.integers(0, 2)), **kw)
    a.use_fused = fused
    a.method = method
    a.integrate()
    yr = np.asarray(r.y[-1], dtype=np.float64)
    return dict(kind="nbody", bodies=nb, method=method, tol=tol, use_fused=fused, steps_ref=len(r) - 1, steps_gpu=len(a) - 1,
                steps_equal=len(r) == len(a), nfev_ref=int(r.nfev), nfev_gpu=int(a.nfev), shape_gpu=list(a.y.shape),
                final_t_err=abs(float(a.t[-1]) - float(r.t[-1])),
                final_rel_err=float(np.max(np.abs(a.y[-1].cpu().numpy() - yr)) / max(1.0, float(np.max(np.abs(yr))))))


def run_event_ensemble_case(ref, de, torch, rng):
    """Events over an ensemble (EventLog) against one reference run per trajectory: an affine plane (located inside the
    fused / stage kernels) or a sphere (search kernels with the callable in between), recorded or terminal."""
    n = 10
    affine = bool(rng.integers(0, 2))
    terminal = bool(rng.integers(0, 2))
    level = float(rng.uniform(20.0, 28.0)) if affine else float(rng.uniform(22.0, 34.0))
    draws = [_draw(rng, "lorenz") for _ in range(n)]
    y0 = np.stack([d[0] for d in draws], axis=1)
    consts = draws[0][1]

    def make(xp):
        if affine:
            def g(t, y, **k):
                return y[2] - level
        else:
            def g(t, y, **k):
                return xp.sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]) - level
        g.is_terminal = terminal
        g.direction = 1
        return [g]
    kw = dict(t=(0.0, 1.5), dt=1e-2, rtol=1e-9, atol=1e-9, constants=consts)
    a = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), ensemble=True, device="cuda:0", **kw)
    a.method = "RK45"
    a.integrate(events=make(torch))
    log = a.events
    t_end = a[-1].t.cpu().numpy()
    y_end = a[-1].y.cpu().numpy()
    missing = extra = missed_terminal = 0
    err = terr = 0.0
    for i in range(n):
        r = ref.OdeSystem(_rhs_factory("lorenz", np), y0=y0[:, i].copy(), dense_output=True, **kw)
        r.method = "RK45"
        r.integrate(events=make(np))
        ev_r = sorted(float(e.t) for e in r.events)
        ev_g = sorted(float(e.t) for e in log.for_trajectory(i))
        # (events the reference found AFTER the time the engine stopped at a terminal crossing are not comparable)
        missing += sum(1 for t in ev_r if t <= float(t_end[i]) + 1e-9 and not any(abs(t - u) <= 1e-9 for u in ev_g))
        extra += max(0, len(ev_g) - len(ev_r))
        if terminal and t_end[i] < float(r.t[-1]) - 1e-9:
            missed_terminal += 1                                 # the reference's root finder passed over a crossing (DESIGN K7)
            continue
        terr = max(terr, abs(float(t_end[i]) - float(r.t[-1])))
        yr = np.asarray(r.y[-1], dtype=np.float64)
        err = max(err, float(np.max(np.abs(y_end[:, i] - yr)) / max(1.0, float(np.max(np.abs(yr))))))
    return dict(kind="events_ens", affine=affine, terminal=terminal, trajectories=n, ref_events_missing_on_gpu=missing,
                gpu_events_beyond_ref=extra, reference_missed_terminal=missed_terminal, final_t_err=terr, final_rel_err=err,
                steps_equal=missing == 0 and terr <= 1e-9)


def run_event_case(ref, de, torch, rng):
    """A terminal and a recorded event on the Lorenz system: event times, states and where the integration stops."""
    y0, consts, _ = _draw(rng, "lorenz")
    level = float(rng.uniform(18.0, 30.0))
    radius = float(rng.uniform(20.0, 35.0))
    terminal = bool(rng.integers(0, 2))

    def make(xp):
        def plane(t, y, **k):
            return y[2] - level

        def sphere(t, y, **k):
            return xp.sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]) - radius
        sphere.is_terminal = terminal
        plane.direction = 1
        return [plane, sphere]
    kw = dict(t=(0.0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, constants=consts, dense_output=True)
    r = ref.OdeSystem(_rhs_factory("lorenz", np), y0=y0.copy(), **kw)
    r.method = "RK45"
    r.integrate(events=make(np))
    a = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0.copy()), device="cuda:0", **kw)
    a.method = "RK45"
    a.integrate(events=make(torch))
    ev_r = sorted(float(e.t) for e in r.events)
    ev_g = sorted(float(e.t) for e in a.events)
    # the reference accepts a located root only if |g(root)| <= 4 eps ABSOLUTE (utilities/optimizer.py:194-197), so it
    # reports a SUBSET of the crossings: every event it reports must be among the engine's
    miss = [t for t in ev_r if not any(abs(t - u) <= 1e-9 for u in ev_g)]
    out = dict(kind="events", terminal=terminal, events_ref=len(ev_r), events_gpu=len(ev_g), steps_ref=len(r) - 1, steps_gpu=len(a) - 1)
    stopped_gpu = "terminated" in str(a.integration_status).lower() or float(a.t[-1]) < 2.0
    stopped_ref = float(r.t[-1]) < 2.0
    if terminal and stopped_gpu and (not stopped_ref or float(a.t[-1]) < float(r.t[-1]) - 1e-9):
        # the engine stopped at a terminal crossing that the reference's root finder did not accept (it went on, or stopped
        # at a later crossing): a documented difference, not compared further
        out.update(reference_missed_terminal_root=True, t_stop_gpu=float(a.t[-1]), t_stop_ref=float(r.t[-1]), steps_equal=True,
                   final_rel_err=0.0)
        return out
    out.update(ref_events_missing_on_gpu=len(miss), final_t_err=float(abs(float(a.t[-1]) - float(r.t[-1]))),
               final_rel_err=float(np.max(np.abs(a.y[-1].cpu().numpy() - np.asarray(r.y[-1]))) / max(1.0, float(np.max(np.abs(r.y[-1]))))),
               steps_equal=len(r) == len(a) and not miss)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--only", default=None, help="comma-separated case kinds")
    args = ap.parse_args()
    import torch
    import desolver_b200 as de
    ref = _import_reference()
    rng = np.random.default_rng(args.seed)
    kinds = (["adaptive"] * 8 + ["fixed"] * 3 + ["symplectic"] * 2 + ["backward"] * 2 + ["dense"] * 3 + ["ensemble"] * 1 + ["events"] * 1 +
             ["continue"] * 2 + ["solve_ivp"] * 3 + ["richardson"] * 1 + ["switch"] * 2 + ["ensemble2"] * 2 + ["shapes"] * 3 + ["events2"] * 1 + ["nbody"] * 2 + ["events_ens"] * 1)
    if args.only:
        kinds = [k for k in kinds if k in args.only.split(",")]
    rows, failures = [], []
    for c in range(args.cases):
        kind = kinds[c % len(kinds)]
        state = rng.bit_generator.state
        try:
            if kind == "ensemble":
                row = run_ensemble_case(ref, de, torch, rng)
            elif kind == "events":
                row = run_event_case(ref, de, torch, rng)
            elif kind == "continue":
                row = run_continue_case(ref, de, torch, rng)
            elif kind == "solve_ivp":
                row = run_solve_ivp_case(ref, de, torch, rng)
            elif kind == "richardson":
                row = run_richardson_case(ref, de, torch, rng)
            elif kind == "switch":
                row = run_switch_case(ref, de, torch, rng)
            elif kind == "ensemble2":
                row = run_ensemble2_case(ref, de, torch, rng)
            elif kind == "shapes":
                row = run_shape_case(ref, de, torch, rng)
            elif kind == "events2":
                row = run_event2_case(ref, de, torch, rng)
            elif kind == "nbody":
                row = run_nbody_case(ref, de, torch, rng)
            elif kind == "events_ens":
                row = run_event_ensemble_case(ref, de, torch, rng)
            else:
                row = run_case(ref, de, torch, rng, kind)
        except Exception as e:                                   # noqa: BLE001 -- a crash IS a finding
            import traceback
            row = dict(kind=kind, error=repr(e)[:300], where=traceback.format_exc()[-700:], steps_equal=False, final_rel_err=float("inf"))
        row["case"] = c
        rows.append(row)
        if args.verbose:
            print(json.dumps(row), file=sys.stderr)
        bar = max(1e-9, 1e-2 * (row.get("tol") or 0.0))          # loose tolerances: step sizes ride on rounding noise
        row["nfev_equal"] = row.get("nfev_ref") == row.get("nfev_gpu")
        row["api_equal"] = not row.get("api_mismatch")
        if "error" in row or not row.get("steps_equal", True) or not (row.get("final_rel_err", 0.0) <= bar) or \
                not row["api_equal"] or (row.get("steps_equal") and not row["nfev_equal"] and row["kind"] != "richardson"):
            # (Richardson wrappers: the reference's basis integrators each evaluate initial_rhs on their first sub-step; the
            # engine's wrapper shares one evaluation -- 2 to 5 evaluations per RUN, states equal to 1e-13)
            failures.append(row)
    ok = [r for r in rows if "error" not in r]
    summary = dict(cases=len(rows), crashed=len(rows) - len(ok), steps_equal=sum(bool(r.get("steps_equal")) for r in ok),
                   final_rel_err_max=max((r["final_rel_err"] for r in ok), default=None),
                   final_rel_err_median=float(np.median([r["final_rel_err"] for r in ok])) if ok else None,
                   max_t_err=max((r.get("max_t_err", 0.0) for r in ok), default=None),
                   dense_rel_err_max=max((r["dense_rel_err"] for r in ok if "dense_rel_err" in r), default=None),
                   reference_missed_terminal_root=sum(bool(r.get("reference_missed_terminal_root")) for r in ok),
                   skipped=sum("skipped" in r for r in ok),
                   nfev_equal=sum(bool(r.get("nfev_equal")) for r in ok if "nfev_ref" in r),
                   nfev_compared=sum("nfev_ref" in r for r in ok),
                   api_equal=sum(bool(r.get("api_equal")) for r in ok if "api_mismatch" in r),
                   api_compared=sum("api_mismatch" in r for r in ok),
                   by_kind={k: dict(cases=sum(r["kind"] == k for r in rows),
                                    steps_equal=sum(bool(r.get("steps_equal")) for r in rows if r["kind"] == k),
                                    final_rel_err_max=max((r["final_rel_err"] for r in ok if r["kind"] == k), default=None))
                            for k in sorted(set(kinds))},
                   seed=args.seed, failures=failures[:80])
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
