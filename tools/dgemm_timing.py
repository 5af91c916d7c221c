#!/usr/bin/env python
"""K3: the library's FP64 tensor-core GEMM (dsb_dgemm, csrc/dgemm.cu) against cuBLAS DGEMM (torch.matmul) on the
linear benchmark shape A (4096 x 4096) @ Y (4096 x 8192), plus agreement of the two results."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from desolver_b200.backend import cuda_backend as cb  # noqa: E402

shapes = [(4096, 8192, 4096), (512, 256, 512), (128, 128, 32), (1024, 8192, 2048)] if len(sys.argv) < 4 else \
    [tuple(int(v) for v in sys.argv[1:4])]
for M, N, K in shapes:
    torch.manual_seed(0)
    A = torch.randn(M, K, dtype=torch.float64, device="cuda")
    B = torch.randn(K, N, dtype=torch.float64, device="cuda")
    C = torch.empty(M, N, dtype=torch.float64, device="cuda")
    ref = A @ B
    assert cb.dgemm(A, B, out=C) is not None
    torch.cuda.synchronize()
    err = float((C - ref).abs().max() / ref.abs().max())

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ours = timed(lambda: cb.dgemm(A, B, out=C))
    lib = timed(lambda: torch.matmul(A, B, out=ref))
    fl = 2.0 * M * N * K
    print(json.dumps(dict(shape=[M, N, K], dsb_dgemm_ms=ours, dsb_dgemm_tflops=fl / ours / 1e9, cublas_ms=lib,
                          cublas_tflops=fl / lib / 1e9, ratio=lib / ours, max_rel_diff=err)))
