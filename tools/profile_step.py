#!/usr/bin/env python
"""Short driver for ncu: a few passes of the headline hot path (2^20 Lorenz trajectories, RK45CK)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--problem", default="lorenz", choices=["lorenz", "vdp"])
ap.add_argument("--method", default="RK45")
ap.add_argument("--strict", action="store_true")
args = ap.parse_args()

if args.problem == "lorenz":
    cfg = de.benchmarks.LORENZ
    y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(args.n)).cuda()
    a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, cfg["tf"]), dt=cfg["dt0"], rtol=1e-9, atol=1e-9, ensemble=True,
                     strict=args.strict)
else:
    cfg = de.benchmarks.VDP
    y0, mu = de.benchmarks.vdp_ensemble(args.n)
    a = de.OdeSystem(de.rhs_plugins.van_der_pol, torch.as_tensor(y0).cuda(), t=(0.0, cfg["tf"]), dt=cfg["dt0"], rtol=1e-9,
                     atol=1e-9, constants=dict(mu=torch.as_tensor(mu).cuda()), ensemble=True, strict=args.strict)
a.method = args.method
for i in range(args.passes):
    a.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a.integrate()
    e1.record()
    torch.cuda.synchronize()
    att = int((a.accepted + a.rejected).sum())
    print("pass %d: %.3f ms, %d attempts, %.3e trajectory-steps/s" % (i, e0.elapsed_time(e1), att, att / e0.elapsed_time(e1) * 1e3))
