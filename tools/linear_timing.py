#!/usr/bin/env python
"""Timing + sampled parity of BASELINE config 2: y' = A y, dense A 4096x4096, 8192 columns, Dormand-Prince 8(7),
per-column step-size control; A @ Y_stage is one FP64 GEMM per stage (the library's own tensor-core kernel dsb_dgemm,
or cuBLAS with --library-gemm), everything else runs in the stage-level kernels of libdesolver_b200.so."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--columns", type=int, default=8192)
ap.add_argument("--check", type=int, default=16, help="columns compared with the CPU oracle")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--library-gemm", action="store_true", help="cuBLAS DGEMM instead of the library's own dsb_dgemm")
args = ap.parse_args()

de.rhs_plugins.linear.use_library_gemm = args.library_gemm
A, Y0 = de.benchmarks.linear_system(args.n, args.columns)
cfg = de.benchmarks.LINEAR
Ad = torch.as_tensor(A).cuda()
a = de.OdeSystem(de.rhs_plugins.linear, torch.as_tensor(Y0).cuda(), t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"],
                 atol=cfg["atol"], constants=dict(A=Ad), ensemble=True)
a.method = "RK87"
all_ms = []
for rep in range(args.reps):
    a.reset()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a.integrate()
    e1.record()
    torch.cuda.synchronize()
    all_ms.append(e0.elapsed_time(e1))
ms = min(all_ms)          # passes vary run to run (allocation / release of ~7 GB of stage arrays and graph pools per pass)
att = (a.accepted + a.rejected)
rounds = a.equ_rhs.nfev // 13
flops_algorithmic = float(att.sum()) * 14 * 2.0 * args.n ** 2          # (S+1) * 2 n^2 per column-step attempt
flops_executed = rounds * 13 * 2.0 * args.n ** 2 * args.columns         # lock-step GEMMs actually issued
out = dict(config="linear n=%d columns=%d RK8713M" % (args.n, args.columns), ms=ms, ms_all=all_ms,
           gemm="cuBLAS (torch.matmul)" if de.rhs_plugins.linear.use_library_gemm else "dsb_dgemm (csrc/dgemm.cu)",
           column_step_attempts=int(att.sum()), attempts_per_column_mean=float(att.double().mean()),
           attempts_per_column_max=int(att.max()), lockstep_rounds=int(rounds),
           column_steps_per_s=float(att.sum()) / ms * 1e3,
           tflops_algorithmic=flops_algorithmic / ms / 1e9, tflops_executed=flops_executed / ms / 1e9)
# GEMM alone, for the roofline denominator of this config
Y = torch.randn(args.n, args.columns, dtype=torch.float64, device="cuda")
torch.matmul(Ad, Y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    torch.matmul(Ad, Y)
e1.record()
torch.cuda.synchronize()
out["dgemm_tflops"] = 5 * 2.0 * args.n ** 2 * args.columns / e0.elapsed_time(e1) / 1e9
if args.check:
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    idx = np.linspace(0, args.columns - 1, args.check).astype(int)
    t0 = time.time()
    ref = orc.erk_ensemble(orc.RHS_LINEAR, Y0[:, idx], cfg["t0"], cfg["tf"], cfg["dt0"], cfg["rtol"], cfg["atol"],
                           *tb.rk_arrays("RK8713MSolver"), shared=A, threads=os.cpu_count())
    got = a[-1].y[:, idx].cpu().numpy()
    err = np.max(np.abs(got - ref["y"]), axis=0) / np.max(np.abs(ref["y"]), axis=0)
    same = (a.accepted[idx].cpu().numpy() == ref["accepted"]) & (a.rejected[idx].cpu().numpy() == ref["rejected"])
    out.update(oracle_columns=int(args.check), oracle_seconds=time.time() - t0, max_rel_err=float(err.max()),
               counts_identical=float(same.mean()))
print(json.dumps(out))
