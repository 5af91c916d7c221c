#!/usr/bin/env python
"""Checks and times the INT8 tensor-core FP64 product (csrc/ozaki_gemm.cu) on the GPU box.

    python tools/ozaki_check.py M N K [--lib path.so] [--time]

Steps: (1) the digit planes reconstruct A and B to 2^-55 of the row / column maximum, (2) the raw INT32 accumulators equal
the exact plane products, (3) C against an exactly rounded reference on sampled entries (rational arithmetic) and against
cuBLAS DGEMM on all entries, (4) timing of split_b + product against cuBLAS and dsb_dgemm."""
import argparse
import ctypes as C
import json
import os
import sys
from fractions import Fraction

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ap = argparse.ArgumentParser()
ap.add_argument("M", type=int)
ap.add_argument("N", type=int)
ap.add_argument("K", type=int)
ap.add_argument("--lib", default=os.path.join(HERE, "..", "desolver_b200", "libdesolver_b200.so"))
ap.add_argument("--time", action="store_true")
ap.add_argument("--no-acc", action="store_true")
ap.add_argument("--only-time", action="store_true")
ap.add_argument("--spread", type=float, default=0.0, help="decades of dynamic range inside rows / columns")
args = ap.parse_args()
M, N, K = args.M, args.N, args.K
lib = C.CDLL(os.path.abspath(args.lib))
vp, ll = C.c_void_p, C.c_longlong
lib.dsb_ozaki_slices.restype = C.c_int
lib.dsb_ozaki_split_a.argtypes = [vp, ll, ll, vp, vp, vp]
lib.dsb_ozaki_split_b.argtypes = [vp, ll, ll, vp, vp, vp, vp]
lib.dsb_ozaki_gemm.argtypes = [vp, vp, vp, vp, vp, ll, ll, ll, C.c_int, vp, vp]
S = lib.dsb_ozaki_slices()
dev = torch.device("cuda")
g = torch.Generator(device="cpu").manual_seed(1)
A = torch.randn(M, K, dtype=torch.float64, generator=g)
B = torch.randn(K, N, dtype=torch.float64, generator=g)
if args.spread > 0:
    A = A * 10.0 ** (args.spread * (torch.rand(M, K, dtype=torch.float64, generator=g) - 0.5))
    B = B * 10.0 ** (args.spread * (torch.rand(K, N, dtype=torch.float64, generator=g) - 0.5))
A, B = A.to(dev), B.to(dev)
a_pl = torch.zeros(S, M, K, dtype=torch.int8, device=dev)
b_pl = torch.zeros(S, N, K, dtype=torch.int8, device=dev)
a_sc = torch.zeros(M, dtype=torch.float64, device=dev)
b_sc = torch.zeros(N, dtype=torch.float64, device=dev)
ws = torch.zeros(N, dtype=torch.int64, device=dev)
Cm = torch.full((M, N), float("nan"), dtype=torch.float64, device=dev)
st = vp(torch.cuda.current_stream().cuda_stream)
sms = torch.cuda.get_device_properties(0).multi_processor_count
out = dict(M=M, N=N, K=K, slices=S)


def check(rc, what):
    if rc != 0:
        print(json.dumps(dict(out, failed=what, rc=rc)))
        sys.exit(1)


check(lib.dsb_ozaki_split_a(A.data_ptr(), M, K, a_pl.data_ptr(), a_sc.data_ptr(), st), "split_a")
check(lib.dsb_ozaki_split_b(B.data_ptr(), K, N, b_pl.data_ptr(), b_sc.data_ptr(), ws.data_ptr(), st), "split_b")
torch.cuda.synchronize()
w = torch.tensor([2.0 ** -(6 + 7 * s) for s in range(S)], dtype=torch.float64, device=dev)
A_rec = (a_pl.double() * w[:, None, None]).sum(0) * a_sc[:, None]
B_rec = ((b_pl.double() * w[:, None, None]).sum(0) * b_sc[:, None]).T
out["split_a_err_rel_rowmax"] = float(((A_rec - A).abs() / a_sc[:, None]).max())
out["split_b_err_rel_colmax"] = float(((B_rec - B).abs() / b_sc[None, :]).max())
out["digits_max"] = [int(a_pl.abs().max()), int(b_pl.abs().max())]
out["scale_ok"] = bool((A.abs().amax(1) < a_sc).all() and (B.abs().amax(0) < b_sc).all())

dbg = None if args.no_acc else torch.zeros(S, M, N, dtype=torch.int32, device=dev)
check(lib.dsb_ozaki_gemm(a_pl.data_ptr(), a_sc.data_ptr(), b_pl.data_ptr(), b_sc.data_ptr(), Cm.data_ptr(), M, N, K, sms,
                         None if dbg is None else dbg.data_ptr(), st), "gemm")
torch.cuda.synchronize()
if dbg is not None:
    ad, bd = a_pl.double(), b_pl.double()
    bad = []
    for d in range(S):
        ref = torch.zeros(M, N, dtype=torch.float64, device=dev)
        for i in range(d + 1):
            ref += ad[i] @ bd[d - i].T
        diff = (dbg[d].double() - ref).abs()
        bad.append(int((diff != 0).sum()))
        if bad[-1] and len(bad) <= 2:
            idx = (diff != 0).nonzero()[:8].tolist()
            out["first_bad_d%d" % d] = [(r, c, int(dbg[d][r, c]), int(ref[r, c])) for r, c in idx]
    out["acc_mismatches_per_weight"] = bad
ref = A @ B
if args.only_time:
    args.time = True
scale = (A.abs() @ B.abs())
out["nan_in_C"] = int(torch.isnan(Cm).sum())
out["vs_cublas_max_err_over_absA_absB"] = float(((Cm - ref).abs() / scale).max())
# exactly rounded reference on sampled entries
rng = np.random.default_rng(0)
samples = 0 if args.only_time else 24
An, Bn, Cn, Rn = A.cpu().numpy(), B.cpu().numpy(), Cm.cpu().numpy(), ref.cpu().numpy()
e_oz, e_cb = 0.0, 0.0
for _ in range(samples):
    i, j = int(rng.integers(M)), int(rng.integers(N))
    exact = sum(Fraction(float(x)) * Fraction(float(y)) for x, y in zip(An[i], Bn[:, j]))
    den = float(sum(abs(Fraction(float(x)) * Fraction(float(y))) for x, y in zip(An[i], Bn[:, j])))
    e_oz = max(e_oz, abs(float(Fraction(float(Cn[i, j])) - exact)) / den)
    e_cb = max(e_cb, abs(float(Fraction(float(Rn[i, j])) - exact)) / den)
out["err_vs_exact_over_sum_abs_products"] = dict(ozaki=e_oz, cublas=e_cb, eps=2.0 ** -53)

if args.time:
    def timed(fn, reps=10):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    out["ms_cublas"] = timed(lambda: torch.matmul(A, B, out=ref))
    out["ms_split_a"] = timed(lambda: lib.dsb_ozaki_split_a(A.data_ptr(), M, K, a_pl.data_ptr(), a_sc.data_ptr(), st))
    out["ms_split_b"] = timed(lambda: lib.dsb_ozaki_split_b(B.data_ptr(), K, N, b_pl.data_ptr(), b_sc.data_ptr(), ws.data_ptr(), st))
    out["ms_gemm"] = timed(lambda: lib.dsb_ozaki_gemm(a_pl.data_ptr(), a_sc.data_ptr(), b_pl.data_ptr(), b_sc.data_ptr(),
                                                      Cm.data_ptr(), M, N, K, sms, None, st))
    out["int8_tops"] = S * (S + 1) / 2 * 2.0 * M * N * K / out["ms_gemm"] / 1e9
    out["fp64_equiv_tflops"] = 2.0 * M * N * K / (out["ms_gemm"] + out["ms_split_b"]) / 1e9
print(json.dumps(out))
