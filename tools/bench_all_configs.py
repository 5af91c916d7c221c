#!/usr/bin/env python
"""Runs every BASELINE.json config on ONE GPU and prints one JSON line per config (kept under profiles/).
Config 1 (the headline) is bench.py's; this driver adds configs 0, 2, 3, 4 with their own bounds:
  0 README oscillator (single system)            -- correctness/latency
  2 linear 4096 x 8192, DP8(7)                   -- FP64 tensor pipe (DGEMM rate)
  3 N-body 64 x 65536, BABs9o7H                  -- FP64 pipe (issue-slot model)
  4 Van der Pol 2^24, RK45CK                     -- HBM (72 B per trajectory-step)"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


# ---- config 0
a = de.OdeSystem(de.rhs_plugins.harmonic_oscillator, np.array([1.0, 0.0]), t=(0, 2 * np.pi), dt=0.01, rtol=1e-9, atol=1e-9,
                 constants=dict(k=1.0, m=1.0))
ms = timed(lambda: (a.reset(), a.integrate()))
print(json.dumps(dict(config=0, what="README harmonic oscillator, RK45 adaptive, single system on the GPU", ms=ms,
                      accepted=int(a.accepted[0]), nfev=a.nfev, y=[float(v) for v in a[-1].y.cpu()],
                      reference=dict(accepted=84, nfev=590, y=[1.0000000023723876, -7.77490676e-11]))))

# ---- config 3
state, m = de.benchmarks.nbody_ensemble(65536, bodies=64, seed=0)
y0 = torch.as_tensor(np.ascontiguousarray(np.moveaxis(state, 1, -1))).cuda()
for method in ("BABS9O7H", "ABAS5O6H"):
    s = de.OdeSystem(de.rhs_plugins.nbody, y0, t=(0.0, 0.1), dt=0.01, constants=dict(m=torch.as_tensor(m), eps2=1e-4, G=1.0), ensemble=True)
    s.method = method
    ms = timed(lambda: (s.reset(), s.integrate()))
    steps = int(s.accepted.sum())
    kicks = int((s.integrator.tableau_intermediate[:, 2] != 0).sum())
    print(json.dumps(dict(config=3, what="N-body 64 bodies x 65536 systems, %s, dt=0.01, 10 steps" % method, ms=ms,
                          system_steps_per_s=steps / ms * 1e3, pair_interactions_per_s=steps * kicks * 4096 / ms * 1e3,
                          bound="fp64 pipe (ncu: 65 % busy = saturated, profiles/README.md issue-slot model)",
                          reference_cpu="126 system-steps/s on one core (BASELINE.md)")))
del s, y0
torch.cuda.empty_cache()

# ---- configs 4 and 2 through their drivers
for cmd in (["tools/vdp_sharded.py", "--check", "1024", "--reps", "3"], ["tools/linear_timing.py", "--check", "8", "--reps", "3"]):
    out = subprocess.run([sys.executable] + cmd, cwd=ROOT, capture_output=True, text=True)
    line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
    print(line)
