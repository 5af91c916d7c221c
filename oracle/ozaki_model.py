"""TEST INFRASTRUCTURE -- a numpy restatement of the FP64-accurate product on INT8 digit planes (csrc/ozaki_gemm.cu), used by
tests/ only (never by the product path).

The reference evaluates the linear benchmark system's `A @ y` with numpy / OpenBLAS inside compute_step
(integrators/components/runge_kutta_methods.py:49-54); the CUDA path may serve it from exact INT8 x INT8 -> INT32 plane
products.  Every step of that scheme is deterministic integer or exactly-rounded FP64 arithmetic, so this model reproduces
the kernel BIT FOR BIT: digit planes, scales, the eight accumulators and the final doubles.

    scale  = 2^e with max|x| = f 2^e, 0.5 <= f < 1        (1 for an all-zero vector)
    r      = x / scale * 64;  q_s = rint(r);  r = (r - q_s) * 128          s = 0 .. 7      (|q| <= 64, exact in FP64)
    acc_d  = sum over i + j = d of  A_i @ B_j^T                             d = 0 .. 7      (exact integers < 2^31)
    C      = (((acc_7 / 128 + acc_6) / 128 + ...) / 128 + acc_0) * scale_a / 4096 * scale_b     (one rounding per addition)
"""
import numpy as np

S = 8


def pow2_above(mx):
    mx = np.asarray(mx, dtype=np.float64)
    _, e = np.frexp(mx)
    out = np.ldexp(1.0, e)
    return np.where((mx > 0) & np.isfinite(mx), out, 1.0)


def split(x):
    """x (R, K) float64: one scale per ROW.  Returns planes (S, R, K) int8 and scale (R,)."""
    x = np.asarray(x, dtype=np.float64)
    scale = pow2_above(np.abs(x).max(axis=1))
    r = x * (1.0 / scale)[:, None] * 64.0
    planes = np.empty((S,) + x.shape, dtype=np.int8)
    for s in range(S):
        n = np.rint(r)
        planes[s] = n.astype(np.int8)
        r = (r - n) * 128.0
    return planes, scale


def accumulators(a_planes, b_planes):
    """acc[d] = sum_{i+j=d} A_i @ B_j^T as exact integers; a_planes (S, M, K), b_planes (S, N, K)."""
    a, b = a_planes.astype(np.int64), b_planes.astype(np.int64)
    acc = np.zeros((S, a.shape[1], b.shape[1]), dtype=np.int64)
    for d in range(S):
        for i in range(d + 1):
            acc[d] += a[i] @ b[d - i].T
    return acc


def combine(acc, a_scale, b_scale):
    s = acc[S - 1].astype(np.float64)
    for d in range(S - 2, -1, -1):
        s = s * (1.0 / 128.0) + acc[d].astype(np.float64)       # s / 128 is exact: the sum rounds once, like the kernel's fma
    return s * (a_scale * (1.0 / 4096.0))[:, None] * b_scale[None, :]


def product(A, B):
    """C = A @ B for A (M, K), B (K, N) the way the INT8 kernel computes it."""
    a_planes, a_scale = split(A)
    b_planes, b_scale = split(np.ascontiguousarray(np.asarray(B, dtype=np.float64).T))
    acc = accumulators(a_planes, b_planes)
    assert np.abs(acc).max() < 2 ** 31
    return combine(acc, a_scale, b_scale), dict(a_planes=a_planes, b_planes=b_planes, a_scale=a_scale, b_scale=b_scale, acc=acc)
