"""GPU tests of the ABI-2 additions to the fused path: kernel-side initialisation (`fresh`, `y_init`), the
statistics counters, chunked launches over one [dim][stride] allocation, and the host_io pipeline of
`OdeSystem` (pinned host y0 in, pinned host results out, uploads / kernels / downloads overlapped).
Every variant must reproduce the plain device-resident run bit for bit: a trajectory's arithmetic never
depends on how the ensemble is cut or where its state came from."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden, rel_inf

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _de():
    import desolver_b200 as de
    return de


def _device_run(y0, tf=2.0, n_reset=0, **kw):
    de = _de()
    a = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0).cuda(), t=(0.0, tf), dt=1e-2, rtol=1e-9, atol=1e-9,
                     ensemble=True, **kw)
    a.method = "RK45"
    a.integrate()
    for _ in range(n_reset):
        a.reset()
        a.integrate()
    torch.cuda.synchronize()
    return a


def test_reset_and_rerun_is_bitwise_and_stats_match_counters():
    de = _de()
    y0 = de.benchmarks.lorenz_ensemble(20000, seed=5)
    a = _device_run(y0)
    b = _device_run(y0, n_reset=2)                 # pristine -> kernel-side init every time
    assert torch.equal(a[-1].y, b[-1].y) and torch.equal(a.accepted, b.accepted) and torch.equal(a.rejected, b.rejected)
    assert torch.equal(a[-1].t, b[-1].t) and torch.equal(a.dt, b.dt)
    st = b.stats
    assert st["attempts"] == int((b.accepted + b.rejected).sum()) and st["accepted"] == int(b.accepted.sum())
    assert st["failed"] == 0 and st["running"] == 0
    # reset() zeroes the evaluation counter (reference :906): the constructor's evaluation is counted only before it
    assert a.nfev == 2 * 20000 + 7 * st["attempts"] and b.nfev == 1 * 20000 + 7 * st["attempts"]
    # a materialised (non-pristine) start takes the in/out path of the kernel: same bits
    c = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0).cuda(), t=(0.0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9,
                     ensemble=True)
    c.method = "RK45"
    assert int(c.accepted.sum()) == 0              # reading a state tensor materialises the state
    c.integrate()
    assert torch.equal(a[-1].y, c[-1].y) and torch.equal(a.accepted, c.accepted)


@pytest.mark.parametrize("mode", ["streamed", "chunked"])
@pytest.mark.parametrize("n,chunks", [(20000, 4), (4099, 3), (1000, 4), (65536, 8), (300000, 16)])
def test_host_io_pipeline_matches_device_run_bitwise(n, chunks, mode):
    de = _de()
    y0 = de.benchmarks.lorenz_ensemble(n, seed=7)
    a = _device_run(y0)
    y0_pinned = torch.from_numpy(y0).pin_memory()
    h = de.OdeSystem(de.rhs_plugins.lorenz, y0_pinned, t=(0.0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True,
                     host_io=True, host_chunks=chunks)
    h.method = "RK45"
    h.host_pipeline = mode         # "streamed": one launch + stream memory operations; "chunked": one launch per chunk
    for rep in range(2):                             # second pass: reset() + integrate() reuses every buffer
        h.integrate()
        assert h.host_pipeline == mode
        assert not h[-1].y.is_cuda and h[-1].y.is_pinned()
        assert np.array_equal(h[-1].y.numpy(), a[-1].y.cpu().numpy())
        assert np.array_equal(h.accepted.numpy(), a.accepted.cpu().numpy())
        assert np.array_equal(h.rejected.numpy(), a.rejected.cpu().numpy())
        assert np.all(h.status.numpy() == 0) and h.success
        assert h.stats["attempts"] == int((a.accepted + a.rejected).sum())
        assert float(h[-1].t.max()) == 2.0 and float(h[-1].t.min()) == 2.0
        h.reset()


def test_host_io_per_trajectory_constants_and_golden():
    de = _de()
    g = load_golden("vdp_rk45.npz")
    y0 = torch.from_numpy(np.ascontiguousarray(g["y0"])).pin_memory()
    h = de.OdeSystem(de.rhs_plugins.van_der_pol, y0, t=(0.0, 10.0), dt=1e-2, rtol=1e-9, atol=1e-9,
                     constants=dict(mu=torch.as_tensor(g["mu"])), ensemble=True, host_io=True, host_chunks=3)
    h.method = "RK45"
    h.integrate()
    acc, rej = h.accepted.numpy(), h.rejected.numpy()
    assert np.mean((acc == g["accepted"]) & (acc + rej == g["attempts"])) >= 0.999
    assert rel_inf(h[-1].y.numpy(), g["y_final"]).max() <= 1e-10


def test_host_io_rejects_what_it_does_not_implement():
    de = _de()
    y0 = torch.from_numpy(de.benchmarks.lorenz_ensemble(256))
    with pytest.raises(NotImplementedError):
        de.OdeSystem(de.rhs_plugins.forced_oscillator, y0[:2], ensemble=True, host_io=True)
    h = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 0.1), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True, host_io=True)
    with pytest.raises(de.exception_types.FailedIntegration):
        h.integrate(callback=lambda s: None)


def test_abi_chunked_launch_with_stride_and_y_init():
    """Straight through the C ABI: two launches over the two halves of one [3][n] allocation (stride = n, y_init
    read-only, fresh = 1) equal one launch over the whole ensemble."""
    de = _de()
    from desolver_b200.backend import cuda_backend as cb
    from desolver_b200.integrators import tableaux as tb
    n = 6000
    y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(n, seed=11)).cuda()
    whole = _device_run(y0.cpu().numpy(), tf=1.0)
    inter, fin, order = tb.rk_arrays("RK45CKSolver")[:3]
    handle = cb.erk_handle(0, de.rhs_plugins.RHS_LORENZ, 3, inter, fin, order)
    y = torch.full((3, n), float("nan"), dtype=torch.float64, device="cuda")
    t = torch.empty(n, dtype=torch.float64, device="cuda")
    dt = torch.empty(n, dtype=torch.float64, device="cuda")
    acc = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    rej = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    st = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    ctr = torch.zeros(16, dtype=torch.int64, device="cuda")
    prm = torch.tensor([10.0, 28.0, 8.0 / 3.0], dtype=torch.float64, device="cuda")
    for c, (lo, hi) in enumerate([(0, 2976), (2976, n)]):
        a = cb.ErkRunArgs()
        a.n, a.stride = hi - lo, n
        a.y, a.y_init = y.data_ptr() + 8 * lo, y0.data_ptr() + 8 * lo
        a.t, a.dt = t.data_ptr() + 8 * lo, dt.data_ptr() + 8 * lo
        for i in range(3):
            a.params[i] = prm.data_ptr() + 8 * i
            a.param_stride[i] = 0
        a.tf, a.rtol, a.atol = 1.0, 1e-9, 1e-9
        a.accepted, a.rejected, a.status = acc.data_ptr() + 4 * lo, rej.data_ptr() + 4 * lo, st.data_ptr() + 4 * lo
        a.max_steps, a.first_call = -1, 1
        a.fresh, a.t_init, a.dt_init = 1, 0.0, 1e-2
        a.counters = ctr.data_ptr() + 64 * c
        handle.run(a, cb.current_stream_ptr(torch.device("cuda:0")))
    torch.cuda.synchronize()
    assert torch.equal(y, whole[-1].y) and torch.equal(acc, whole.accepted) and torch.equal(rej, whole.rejected)
    assert int(st.abs().sum()) == 0 and torch.equal(t, whole[-1].t)
    c = ctr.cpu().reshape(2, 8)
    assert int(c[:, 2].sum()) == int((acc + rej).sum()) and int(c[:, 3].sum()) == int(acc.sum())
    assert int(c[0, 0]) >= 2976 and int(c[:, 4].sum()) == 0
    # stride smaller than n is refused
    bad = cb.ErkRunArgs()
    bad.n, bad.stride = 10, 5
    bad.y, bad.t, bad.dt, bad.counters = y.data_ptr(), t.data_ptr(), dt.data_ptr(), ctr.data_ptr()
    assert cb.lib().dsb_erk_run(handle._h, C.byref(bad), None) == cb.DSB_ERR_BAD_ARGUMENT


@pytest.mark.parametrize("mode", ["streamed", "chunked"])
@pytest.mark.parametrize("n", [1, 31, 33, 1000])
def test_host_io_tiny_ensembles_and_pageable_memory(n, mode):
    """Sizes below one warp / one chunk, and y0 in ordinary (pageable) host memory: same bits as the device run."""
    de = _de()
    y0 = de.benchmarks.lorenz_ensemble(n, seed=3)
    ref = _device_run(y0, tf=0.5)
    h = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(0.0, 0.5), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True,
                     host_io=True)
    h.method = "RK45"
    h.host_pipeline = mode
    h.integrate()
    assert torch.equal(h[-1].y, ref[-1].y.cpu()) and torch.equal(h.accepted, ref.accepted.cpu())
