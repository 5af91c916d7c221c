#!/usr/bin/env python
"""Compile-time parameter sweep of the FP64 tensor-core GEMM (csrc/dgemm.cu): builds one .so per variant HERE
(`python tools/dgemm_sweep.py build`), times them on the GPU box (`python tools/dgemm_sweep.py`) against cuBLAS."""
import ctypes as C
import glob
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "dgemm_variants")
VARIANTS = {
    "bk16_s3_c2": "",
    "bk16_s3_c2_ldb4": "-DDSB_GEMM_LDB_PAD=4",
    "bk16_s3_c2_ldb12": "-DDSB_GEMM_LDB_PAD=12",
    "bk16_s4_c2_ldb4": "-DDSB_GEMM_LDB_PAD=4 -DDSB_GEMM_STAGES=4",
    "bk16_s3_c2_ldb4_fdb": "-DDSB_GEMM_LDB_PAD=4 -DDSB_GEMM_FRAG_DB=1",
    "bk16_s3_c2_g16": "-DDSB_GEMM_GROUP_M=16",
    "bk16_s3_c2_fdb": "-DDSB_GEMM_FRAG_DB=1",
    "bk16_s4_c2": "-DDSB_GEMM_STAGES=4",
}
if len(sys.argv) > 1 and sys.argv[1] == "build":
    os.makedirs(OUT, exist_ok=True)
    for name, flags in VARIANTS.items():
        cmd = "nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr " \
              "-diag-suppress 128 %s -shared -o %s/%s.so %s/dgemm_variant.cu -lcudart" % (flags, OUT, name, HERE)
        r = subprocess.run(cmd, shell=True, capture_output=True, text=True)
        print(name, "ok" if r.returncode == 0 else r.stderr[-400:])
    sys.exit(0)

import torch  # noqa: E402
M, N, K = 4096, 8192, 4096
torch.manual_seed(0)
A = torch.randn(M, K, dtype=torch.float64, device="cuda")
B = torch.randn(K, N, dtype=torch.float64, device="cuda")
Cm = torch.empty(M, N, dtype=torch.float64, device="cuda")
ref = A @ B
sms = torch.cuda.get_device_properties(0).multi_processor_count


def timed(fn, reps=10):
    fn()
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = dict(cublas_ms=timed(lambda: torch.matmul(A, B, out=ref)))
for path in sorted(glob.glob(os.path.join(OUT, "*.so"))):
    lib = C.CDLL(path)
    lib.variant_dgemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    Cm.zero_()
    rc = lib.variant_dgemm(A.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, K, sms, st)
    torch.cuda.synchronize()
    name = os.path.basename(path)[:-3]
    if rc != 0:
        out[name] = "rc=%d" % rc
        continue
    err = float((Cm - ref).abs().max() / ref.abs().max())
    ms = timed(lambda: lib.variant_dgemm(A.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, K, sms, st))
    out[name] = dict(ms=ms, tflops=2.0 * M * N * K / ms / 1e9, max_rel_diff=err)
    # one CTA per tile (the hardware block scheduler balances the tail instead of the static round-robin)
    out[name]["ms_grid_per_tile"] = timed(lambda: lib.variant_dgemm(A.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, K, 1 << 20, st))
print(json.dumps(out, indent=1))
