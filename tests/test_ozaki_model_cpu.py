"""CPU tests of the digit-plane product's numpy model (oracle/ozaki_model.py), the checker of csrc/ozaki_gemm.cu: the model
itself is pinned against exact rational arithmetic here; tests/test_gpu_int8_gemm.py compares the kernel with it bit for bit."""
from fractions import Fraction

import numpy as np

from oracle import ozaki_model as om

EPS = 2.0 ** -53


def _exact_err(A, B, C, entries):
    worst = 0.0
    for i, j in entries:
        terms = [Fraction(float(x)) * Fraction(float(y)) for x, y in zip(A[i], B[:, j])]
        den = float(sum(abs(t) for t in terms))
        worst = max(worst, abs(float(Fraction(float(C[i, j])) - sum(terms))) / den)
    return worst


def test_planes_are_an_error_free_split():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((16, 96)) * 10.0 ** rng.uniform(-3, 3, (16, 96))
    x[3] = 0.0
    x[4] = 1.0
    planes, scale = om.split(x)
    assert planes.dtype == np.int8 and np.abs(planes.astype(int)).max() <= 64
    assert scale[3] == 1.0 and scale[4] == 2.0
    assert np.all(np.abs(x).max(1)[np.arange(16) != 3] < scale[np.arange(16) != 3])
    w = np.array([2.0 ** -(6 + 7 * s) for s in range(om.S)])
    rec = np.tensordot(w, planes.astype(np.float64), axes=1) * scale[:, None]
    assert np.max(np.abs(rec - x) / scale[:, None]) <= 2.0 ** -56
    big = np.abs(x) >= scale[:, None] / 4                      # elements within two bits of the row maximum are exact
    assert np.array_equal(rec[big], x[big])


def test_product_is_within_half_an_ulp_of_sum_abs_products():
    rng = np.random.default_rng(1)
    A, B = rng.standard_normal((24, 512)), rng.standard_normal((512, 16))
    C, parts = om.product(A, B)
    assert np.abs(parts["acc"]).max() < 2 ** 27
    entries = [(int(rng.integers(24)), int(rng.integers(16))) for _ in range(12)]
    assert _exact_err(A, B, C, entries) <= EPS
    assert _exact_err(A, B, A @ B, entries) <= 64 * EPS      # numpy's own product, for scale
    assert np.max(np.abs(C - A @ B) / (np.abs(A) @ np.abs(B))) <= 8 * EPS


def test_accumulators_cannot_overflow_at_the_largest_supported_k():
    # worst case: every digit +-64, K = 32768 (the kernel's limit), 8 plane products in the heaviest accumulator
    assert 64 * 64 * 32768 * 8 < 2 ** 31
    a = np.full((om.S, 2, 4096), 64, dtype=np.int8)
    acc = om.accumulators(a, a)
    assert acc.max() == 64 * 64 * 4096 * 8 and acc.max() < 2 ** 27 + 1
