#!/usr/bin/env python
"""The headline ensemble through an ORDINARY torch callable (stage-level kernels + CUDA-graph replay + compaction) next to
the fused plugin: what "arbitrary torch RHS callables between fused stage kernels" costs at scale.

    python tools/callable_timing.py [log2_n=20] > profiles/round2_callable_vs_fused.json
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402


def lorenz(t, y, sigma, rho, beta):
    return torch.stack([sigma * (y[1] - y[0]), y[0] * (rho - y[2]) - y[1], y[0] * y[1] - beta * y[2]])


def run(rhs, y0, reps=3, **extra):
    cfg = de.benchmarks.LORENZ
    out = []
    for _ in range(reps):
        a = de.OdeSystem(rhs, y0, t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"], ensemble=True,
                         constants=dict(sigma=cfg["sigma"], rho=cfg["rho"], beta=cfg["beta"]), **extra)
        a.method = "RK45"
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.integrate()
        torch.cuda.synchronize()
        out.append((time.perf_counter() - t0) * 1e3)
    att = int((a.accepted + a.rejected).sum())
    return a, min(out), out, att


def rossler_lines(n):
    """The same comparison for a right-hand side of the CALLER: Roessler as an untagged torch callable, and as the fused
    device plug-in `rhs_plugins.register_device_rhs` builds from its CUDA body (desolver_b200.benchmarks.rossler_plugin)."""
    import numpy as np
    rng = np.random.default_rng(3)
    y0 = torch.as_tensor(rng.uniform([-8, -8, 0], [8, 8, 12], (n, 3)).T.copy()).cuda()
    kw = dict(t=(0.0, 10.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)

    def untagged(t, y, a=0.2, b=0.2, c=5.7):
        return torch.stack([-y[1] - y[2], y[0] + a * y[1], b + y[2] * (y[0] - c)])
    out = {}
    for key, rhs in (("fused_plugin", de.benchmarks.rossler_plugin()), ("torch_callable", untagged)):
        best = 1e30
        for _ in range(3):
            a = de.OdeSystem(rhs, y0, **kw)
            a.method = "RK45"
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            a.integrate()
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) * 1e3)
        out[key] = (a, best)
    f, c = out["fused_plugin"][0], out["torch_callable"][0]
    att = int((f.accepted + f.rejected).sum())
    return dict(workload="roessler_rk45ck_2^%d, t in [0,10], tol 1e-9" % (n.bit_length() - 1), trajectory_steps=att,
                user_plugin_ms=out["fused_plugin"][1], torch_callable_ms=out["torch_callable"][1],
                user_plugin_trajectory_steps_per_s=att / (out["fused_plugin"][1] * 1e-3),
                hbm_roofline_frac_80B=att * 80.0 / (out["fused_plugin"][1] * 1e-3) / 6538.3e9,
                counts_identical=bool(torch.equal(f.accepted, c.accepted) and torch.equal(f.rejected, c.rejected)),
                max_rel_state_difference=float(((f[-1].y - c[-1].y).abs().amax(0) / c[-1].y.abs().amax(0).clamp_min(1.0)).max()))


if __name__ == "__main__":
    n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
    if len(sys.argv) > 2 and sys.argv[2] == "rossler":
        print(json.dumps(rossler_lines(n), indent=1))
        sys.exit(0)
    y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(n)).cuda()
    f, ms_f, all_f, att = run(de.rhs_plugins.lorenz, y0)
    c, ms_c, all_c, att_c = run(lorenz, y0)
    same = bool(torch.equal(f.accepted, c.accepted) and torch.equal(f.rejected, c.rejected))
    err = float(((f[-1].y - c[-1].y).abs().amax(0) / f[-1].y.abs().amax(0)).max())
    rounds = int((c.accepted + c.rejected).max())
    print(json.dumps(dict(workload="lorenz63_rk45ck_2^%d" % (n.bit_length() - 1), trajectory_steps=att,
                          fused_plugin_ms=ms_f, fused_all=all_f, torch_callable_ms=ms_c, callable_all=all_c,
                          callable_trajectory_steps_per_s=att_c / (ms_c * 1e-3), lock_step_rounds=rounds,
                          counts_identical=same, max_rel_state_difference=err,
                          note="callable: 6 stage kernels + the caller's ~12 elementwise torch kernels per stage, replayed as one CUDA "
                               "graph per lock-step round, working set compacted to the active trajectories"), indent=1))
