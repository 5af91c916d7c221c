#!/usr/bin/env python
"""Short driver for ncu: the stage-level kernels (K2) and the FP64 GEMM of the linear config
(y' = A y, n x n, `columns` columns, Dormand-Prince 8(7)); a few attempts only."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--columns", type=int, default=8192)
ap.add_argument("--tf", type=float, default=0.05)
args = ap.parse_args()
A, Y0 = de.benchmarks.linear_system(args.n, args.columns)
a = de.OdeSystem(de.rhs_plugins.linear, torch.as_tensor(Y0).cuda(), t=(0.0, args.tf), dt=1e-2, rtol=1e-9, atol=1e-9,
                 constants=dict(A=torch.as_tensor(A).cuda()), ensemble=True)
a.method = "RK87"
a.integrate()
torch.cuda.synchronize()
print("attempt rounds", a.equ_rhs.nfev // 13, "accepted", int(a.accepted.sum()))
