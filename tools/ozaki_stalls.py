#!/usr/bin/env python
"""Where the INT8 product's pipeline waits (DSB_OZAKI_MODE=4 instrumentation of csrc/ozaki_gemm.cu): per CTA the cycles the
MMA lane spent waiting for tiles / for the epilogue, and the producers for free slots."""
import ctypes as C
import json
import os
import sys

import torch

os.environ["DSB_OZAKI_MODE"] = "4"
os.environ["DSB_OZAKI_NO_PAIR"] = "1"          # the statistics are implemented in the single-CTA kernel
lib = C.CDLL(os.path.abspath(sys.argv[1]))
M, N, K = 4096, 8192, 4096
vp, ll = C.c_void_p, C.c_longlong
lib.dsb_ozaki_split_a.argtypes = [vp, ll, ll, vp, vp, vp]
lib.dsb_ozaki_split_b.argtypes = [vp, ll, ll, vp, vp, vp, vp]
lib.dsb_ozaki_gemm.argtypes = [vp, vp, vp, vp, vp, ll, ll, ll, C.c_int, vp, vp]
dev = "cuda"
A = torch.randn(M, K, dtype=torch.float64, device=dev)
B = torch.randn(K, N, dtype=torch.float64, device=dev)
S = 8
a_pl = torch.zeros(S, M, K, dtype=torch.int8, device=dev)
b_pl = torch.zeros(S, N, K, dtype=torch.int8, device=dev)
a_sc, b_sc = torch.zeros(M, dtype=torch.float64, device=dev), torch.zeros(N, dtype=torch.float64, device=dev)
ws = torch.zeros(N, dtype=torch.int64, device=dev)
Cm = torch.zeros(M, N, dtype=torch.float64, device=dev)
sms = torch.cuda.get_device_properties(0).multi_processor_count
lib.dsb_ozaki_split_a(A.data_ptr(), M, K, a_pl.data_ptr(), a_sc.data_ptr(), None)
lib.dsb_ozaki_split_b(B.data_ptr(), K, N, b_pl.data_ptr(), b_sc.data_ptr(), ws.data_ptr(), None)
for _ in range(3):
    st = torch.zeros(sms * 5, dtype=torch.int64, device=dev)
    rc = lib.dsb_ozaki_gemm(a_pl.data_ptr(), a_sc.data_ptr(), b_pl.data_ptr(), b_sc.data_ptr(), Cm.data_ptr(), M, N, K, sms,
                            st.data_ptr(), None)
    torch.cuda.synchronize()
s = st[:sms * 4].view(sms, 4).double()
w = st[sms * 4:].double()
print(json.dumps(dict(rc=rc, cycles_total=float(s[:, 0].mean()), mma_wait_tiles=float(w.mean()), mma_wait_epilogue=float(s[:, 3].mean()),
                      producer_a_wait_slots=float(s[:, 1].mean()), producer_b_wait_slots=float(s[:, 2].mean()),
                      frac_mma_waiting_for_tiles=float((w / s[:, 0]).mean()))))
