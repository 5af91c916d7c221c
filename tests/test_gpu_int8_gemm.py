"""GPU tests of the FP64-accurate product on the INT8 tensor cores (dsb_ozaki_*, csrc/ozaki_gemm.cu), the second kernel
behind the `A @ y` of the reference's linear benchmark system (integrators/components/runge_kutta_methods.py:49-54 calls the
user's right-hand side; rhs_plugins.linear serves it).

The product is assembled from exact pieces, so it is tested piece by piece: (1) the digit planes reconstruct the operands,
(2) the raw INT32 accumulators equal the exact plane products -- bit-exact, (3) the FP64 result against an exactly rounded
reference (rational arithmetic) on sampled entries and against an FP64 GEMM on all entries, (4) the edge cases of the
scaling (zero rows, wide dynamic range, powers of two), (5) the whole linear configuration against the CPU oracle."""
import os
from fractions import Fraction

import numpy as np
import pytest

from conftest import rel_inf

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

EPS = 2.0 ** -53


def _cb():
    from desolver_b200.backend import cuda_backend as cb
    return cb


def _operands(M, N, K, seed=0, spread=0.0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, dtype=torch.float64, generator=g)
    B = torch.randn(K, N, dtype=torch.float64, generator=g)
    if spread:
        A = A * 10.0 ** (spread * (torch.rand(M, K, dtype=torch.float64, generator=g) - 0.5))
        B = B * 10.0 ** (spread * (torch.rand(K, N, dtype=torch.float64, generator=g) - 0.5))
    return A.cuda(), B.cuda()


def _planes_value(planes, scale):
    S = planes.shape[0]
    w = torch.tensor([2.0 ** -(6 + 7 * s) for s in range(S)], dtype=torch.float64, device=planes.device)
    return (planes.double() * w[:, None, None]).sum(0) * scale[:, None]


def _exact_errors(A, B, C, n=12, seed=0):
    """max over sampled entries of |C_ij - exact| / sum_k |a_ik b_kj|, exact in rational arithmetic"""
    rng = np.random.default_rng(seed)
    An, Bn, Cn = A.cpu().numpy(), B.cpu().numpy(), C.cpu().numpy()
    worst = 0.0
    for _ in range(n):
        i, j = int(rng.integers(An.shape[0])), int(rng.integers(Bn.shape[1]))
        terms = [Fraction(float(x)) * Fraction(float(y)) for x, y in zip(An[i], Bn[:, j])]
        den = float(sum(abs(t) for t in terms))
        worst = max(worst, abs(float(Fraction(float(Cn[i, j])) - sum(terms))) / den)
    return worst


@pytest.mark.parametrize("shape", [(128, 64, 128), (128, 128, 256), (256, 192, 384), (384, 64, 1024), (512, 1024, 512)])
def test_planes_and_accumulators_are_exact(shape):
    cb = _cb()
    M, N, K = shape
    A, B = _operands(M, N, K, seed=M + N + K)
    prod = cb.Int8Product(A)
    assert prod.S == 8 and prod.supports(B)
    dbg = torch.zeros(prod.S, M, N, dtype=torch.int32, device="cuda")
    C = prod(B, dbg_acc=dbg)
    torch.cuda.synchronize()
    a_pl, b_pl = prod.a_planes, prod.b_planes.view(prod.S, N, K)
    # (1) digits are 7-bit signed, the scale is the power of two above the row / column maximum, 55 bits are kept
    assert int(a_pl.abs().max()) <= 64 and int(b_pl.abs().max()) <= 64
    assert bool((A.abs().amax(1) < prod.a_scale).all()) and bool((A.abs().amax(1) >= prod.a_scale / 2).all())
    assert bool((B.abs().amax(0) < prod.b_scale).all()) and bool((B.abs().amax(0) >= prod.b_scale / 2).all())
    assert float(((_planes_value(a_pl, prod.a_scale) - A).abs() / prod.a_scale[:, None]).max()) <= 2.0 ** -56
    assert float(((_planes_value(b_pl, prod.b_scale).T - B).abs() / prod.b_scale[None, :]).max()) <= 2.0 ** -56
    # (2) accumulator d = sum over i + j = d of A_i B_j^T, exactly (integers below 2^27: exact in float64 too)
    ad, bd = a_pl.double(), b_pl.double()
    for d in range(prod.S):
        ref = sum(ad[i] @ bd[d - i].T for i in range(d + 1))
        assert torch.equal(dbg[d].double(), ref), "weight %d" % d
    # (3) the FP64 result
    assert _exact_errors(A, B, C, n=6) <= EPS
    ref = A @ B
    assert float(((C - ref).abs() / (A.abs() @ B.abs())).max()) <= 8 * EPS


@pytest.mark.parametrize("shape", [(128, 64, 128), (256, 128, 512), (384, 192, 256)])
def test_kernel_equals_the_numpy_model_bit_for_bit(shape):
    """oracle/ozaki_model.py restates the scheme in numpy (rint digits, exact int64 plane products, Horner in float64): digit
    planes, scales, accumulators and the final doubles of the kernel are the model's, bit for bit -- on single CTAs
    (M = 128, 384) and on CTA pairs (M = 256)."""
    from oracle import ozaki_model as om
    cb = _cb()
    M, N, K = shape
    A, B = _operands(M, N, K, seed=5 * M + N, spread=2.0)
    prod = cb.Int8Product(A)
    dbg = torch.zeros(prod.S, M, N, dtype=torch.int32, device="cuda")
    C = prod(B, dbg_acc=dbg)
    Cm, parts = om.product(A.cpu().numpy(), B.cpu().numpy())
    assert np.array_equal(prod.a_planes.cpu().numpy(), parts["a_planes"]) and np.array_equal(prod.a_scale.cpu().numpy(), parts["a_scale"])
    assert np.array_equal(prod.b_planes.view(prod.S, N, K).cpu().numpy(), parts["b_planes"])
    assert np.array_equal(prod.b_scale.cpu().numpy(), parts["b_scale"])
    assert np.array_equal(dbg.cpu().numpy().astype(np.int64), parts["acc"])
    assert np.array_equal(C.cpu().numpy(), Cm)


def test_closer_to_the_exact_product_than_an_fp64_gemm():
    """K = 4096 (the benchmark's contraction length): the error against the exactly rounded product stays below half an
    ulp of sum |a b| -- the left-to-right FP64 GEMM is allowed K times that."""
    cb = _cb()
    A, B = _operands(256, 128, 4096, seed=7)
    C = cb.Int8Product(A)(B)
    e_int8 = _exact_errors(A, B, C, n=16)
    e_fp64 = _exact_errors(A, B, A @ B, n=16)
    assert e_int8 <= EPS
    assert e_int8 <= e_fp64


def test_dynamic_range_inside_rows_and_columns():
    """Elements 6 decades below the largest of their row / column lose their low bits (the planes hold 55 bits below the
    row maximum): the error bound is relative to |row max| |column max|, like a normwise GEMM bound."""
    cb = _cb()
    A, B = _operands(256, 128, 512, seed=3, spread=6.0)
    prod = cb.Int8Product(A)
    C = prod(B)
    ref = A @ B
    bound = prod.a_scale[:, None] * prod.b_scale[None, :] * A.shape[1]
    assert float(((C - ref).abs() / bound).max()) <= 2.0 ** -50
    # and the usual componentwise bound still holds to a few ulps here, because every row has large elements
    assert float(((C - ref).abs() / (A.abs() @ B.abs())).max()) <= 1e-13


def test_zero_rows_columns_and_powers_of_two():
    cb = _cb()
    A, B = _operands(128, 128, 128, seed=5)
    A[3] = 0.0
    B[:, 7] = 0.0
    A[5] = 0.0
    A[5, 9] = 2.0 ** -300                               # tiny row: the scale follows it
    A[6] = 1.0                                          # exactly a power of two: scale must be strictly above
    B[:, 8] = -0.5
    prod = cb.Int8Product(A)
    C = prod(B)
    ref = A @ B
    assert torch.equal(C[3], torch.zeros_like(C[3])) and torch.equal(C[:, 7], torch.zeros_like(C[:, 7]))
    assert float(prod.a_scale[6]) == 2.0 and float(prod.b_scale[8]) == 1.0
    # one product per entry of row 5: exact up to the 55 bits the planes of B keep below each column maximum
    assert float(((C[5] - A[5, 9] * B[9]).abs() / (A[5, 9] * prod.b_scale)).max()) <= 2.0 ** -54
    rel = (C - ref).abs() / (A.abs() @ B.abs() + 1e-300)
    rel[5] = 0.0                                        # row 5 is bounded by the column maxima (above), not componentwise
    assert float(rel.max()) <= 8 * EPS


def test_planes_follow_in_place_updates_of_the_matrix():
    cb = _cb()
    A, B = _operands(128, 64, 128, seed=11)
    prod = cb.Int8Product(A)
    C1 = prod(B).clone()
    A.mul_(3.0)
    C2 = prod(B)
    assert float(((C2 - 3.0 * C1).abs() / (A.abs() @ B.abs())).max()) <= 4 * EPS


def test_matrix_modified_between_integrations_is_seen_by_replayed_graphs():
    """The stage-level path replays its attempts as CUDA graphs; the digit planes of A are cut inside the capture, so an
    in-place update of A between two integrate() calls reaches the replays (as it does with the FP64 kernels)."""
    import desolver_b200 as de
    lin = de.rhs_plugins.linear
    A, Y0 = de.benchmarks.linear_system(256, 256)
    Ad = torch.as_tensor(A).cuda()

    def run(a=None):
        if a is None:
            a = de.OdeSystem(lin, torch.as_tensor(Y0).cuda(), t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, constants=dict(A=Ad),
                             ensemble=True)
            a.method = "RK87"
        else:
            a.reset()
        a.integrate()
        return a
    lin.gemm = "int8"
    try:
        a = run()
        run(a)                                           # replays
        y1 = a[-1].y.clone()
        Ad.mul_(0.5)                                     # y' = (A / 2) y: y(1) of the new system = y(1/2) of the old one
        run(a)
        y2 = a[-1].y.clone()
        fresh = run()
    finally:
        lin.gemm = "auto"
        Ad.mul_(2.0)
    assert float((y2 - y1).abs().max()) > 1e-3           # the update was seen
    assert float((y2 - fresh[-1].y).abs().max()) <= 1e-10 * float(y2.abs().max())


@pytest.mark.parametrize("shape", [(128, 96, 128), (130, 70, 200), (1, 1, 1), (257, 64, 129), (1000, 777, 1500)])
def test_shapes_outside_the_tiling_are_zero_padded_exactly(shape):
    """M, K not multiples of 128 / N not a multiple of 64: operands are zero-padded to the tiling in staging buffers.  A zero
    changes no row / column maximum and no plane product, so the result is STILL the numpy model's, bit for bit, on the
    unpadded shapes."""
    from oracle import ozaki_model as om
    cb = _cb()
    M, N, K = shape
    A, B = _operands(M, N, K, seed=M + 3 * N + K, spread=1.0)
    prod = cb.Int8Product(A)
    assert prod.supports(B) and prod.shape_ok()
    C = prod(B)
    assert tuple(C.shape) == (M, N)
    Cm, _ = om.product(A.cpu().numpy(), B.cpu().numpy())
    assert np.array_equal(C.cpu().numpy(), Cm)
    ref = A @ B
    assert float((C - ref).abs().max()) <= 1e-12 * max(1.0, float(ref.abs().max()))
    out = torch.full((M, N), float("nan"), dtype=torch.float64, device="cuda")
    assert prod(B, out=out) is out and torch.equal(out, C)                        # caller-owned result, second call
    narrower = B[:, : max(1, N // 2)].contiguous()
    assert np.array_equal(prod(narrower).cpu().numpy(), om.product(A.cpu().numpy(), narrower.cpu().numpy())[0])


def test_kernel_arguments_outside_the_tiling_are_refused():
    cb = _cb()
    assert not cb.Int8Product(torch.zeros(64, 40000, dtype=torch.float64, device="cuda")).shape_ok()   # K > 32768
    L = cb.lib()
    assert L.dsb_ozaki_gemm(None, None, None, None, None, 128, 96, 64, 1, None, None) == cb.DSB_ERR_BAD_ARGUMENT
    assert L.dsb_ozaki_gemm(None, None, None, None, None, 128, 64, 65536, 1, None, None) == cb.DSB_ERR_BAD_ARGUMENT
    assert L.dsb_ozaki_split_a(None, 128, 64, None, None, None) == cb.DSB_ERR_BAD_ARGUMENT
    assert b"dsb_ozaki" in cb.lib().dsb_last_error()


def test_linear_plugin_on_the_int8_kernel_matches_the_oracle():
    """y' = A y through `linear.gemm = "int8"`: same accepted / rejected counts as the DMMA kernel on every column, states
    within the parity bar of the CPU oracle on sampled columns; the products really ran on the INT8 kernel."""
    import desolver_b200 as de
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    lin = de.rhs_plugins.linear
    A, Y0 = de.benchmarks.linear_system(256, 256)
    Ad = torch.as_tensor(A).cuda()

    def run(mode):
        lin.gemm = mode
        try:
            a = de.OdeSystem(lin, torch.as_tensor(Y0).cuda(), t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, constants=dict(A=Ad),
                             ensemble=True)
            a.method = "RK87"
            a.integrate()
        finally:
            lin.gemm = "auto"
        return a
    before = lin.calls_int8
    a = run("int8")
    assert lin.calls_int8 > before
    b = run("dmma")
    # two different (both valid) roundings of A @ y: the trajectories agree to the integration tolerance, not to the bit
    assert float((a[-1].y - b[-1].y).abs().max()) <= 1e-10 * float(b[-1].y.abs().max())
    # RK8(7) at dt0 = 1e-2 starts from error estimates of pure rounding noise: an accept / reject decision can flip with
    # the rounding of the product (tests/test_gpu_stage.py::test_linear_plugin_runs_on_the_dgemm_kernel); the step COUNTS
    # of the benchmark configuration are compared with the oracle in tools/config_runs.py (identical on the sample)
    assert float((a.accepted == b.accepted).double().mean()) >= 0.75
    idx = np.arange(0, 256, 16)
    ref = orc.erk_ensemble(orc.RHS_LINEAR, Y0[:, idx], 0.0, 1.0, 1e-2, 1e-9, 1e-9, *tb.rk_arrays("RK8713MSolver"), shared=A,
                           threads=min(4, os.cpu_count()))
    assert rel_inf(a[-1].y[:, idx].cpu().numpy(), ref["y"]).max() <= 1e-10
    assert np.mean(a.accepted[idx].cpu().numpy() == ref["accepted"]) >= 0.75
