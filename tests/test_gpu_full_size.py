"""The BASELINE.json configurations at their FULL sizes, through size-independent properties (the per-trajectory
oracle cannot cover 2^24 trajectories): bitwise determinism, independence of a trajectory from the rest of the
ensemble (a subset integrated on its own gives the same bits), conservation laws, events as pure observers -- plus the
oracle on a sample.  Config 1 (Lorenz 2^20) lives in test_gpu_erk.py::test_deterministic_and_order_independent and
test_gpu_host_io.py."""
import os

import numpy as np
import pytest

from conftest import rel_inf

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _de():
    import desolver_b200 as de
    return de


def test_config4_van_der_pol_2pow24():
    """2^24 trajectories, per-trajectory mu in [0.1, 10], RK45CK, t in [0, 10]."""
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    de = _de()
    n = 1 << 24
    y0, mu = de.benchmarks.vdp_ensemble(n)
    y0d, mud = torch.as_tensor(y0).cuda(), torch.as_tensor(mu).cuda()

    def run(y, m):
        a = de.OdeSystem(de.rhs_plugins.van_der_pol, y, t=(0.0, 10.0), dt=1e-2, rtol=1e-9, atol=1e-9, constants=dict(mu=m),
                         ensemble=True)
        a.method = "RK45"
        a.integrate()
        return a
    a = run(y0d, mud)
    assert a.success and a.stats["failed"] == 0 and a.stats["running"] == 0 and bool((a[-1].t == 10.0).all())
    assert a.stats["attempts"] == int((a.accepted.long() + a.rejected.long()).sum())
    ya, acc_a, rej_a = a[-1].y.clone(), a.accepted.clone(), a.rejected.clone()
    a.reset()
    a.integrate()
    assert torch.equal(ya, a[-1].y) and torch.equal(acc_a, a.accepted) and torch.equal(rej_a, a.rejected)     # determinism
    # a strided subset integrated on its own: the same bits (no trajectory sees its neighbours)
    sub = slice(5, n, 4099)
    b = run(y0d[:, sub].contiguous(), mud[sub].contiguous())
    assert torch.equal(b[-1].y, ya[:, sub]) and torch.equal(b.accepted, acc_a[sub])
    # the oracle on a sample
    idx = np.arange(0, n, n // 1024)[:1024]
    ref = orc.erk_ensemble(orc.RHS_VDP, y0[:, idx], 0.0, 10.0, 1e-2, 1e-9, 1e-9, *tb.rk_arrays("RK45CKSolver"),
                           params=mu[None, idx], threads=os.cpu_count())
    same = (acc_a[idx].cpu().numpy() == ref["accepted"]) & (rej_a[idx].cpu().numpy() == ref["rejected"])
    assert same.mean() >= 0.999 and rel_inf(ya[:, idx].cpu().numpy(), ref["y"]).max() <= 1e-10


def test_config3_nbody_65536_systems():
    """65536 systems x 64 bodies, BABs9o7H, dt = 0.01, 10 steps."""
    de = _de()
    B, nb = 65536, 64
    state, m = de.benchmarks.nbody_ensemble(B, bodies=nb, seed=0)
    y0 = torch.as_tensor(np.ascontiguousarray(np.moveaxis(state, 1, -1))).cuda()        # (2, nb, 3, B)
    consts = dict(m=torch.as_tensor(m), eps2=1e-4, G=1.0)

    def run(y):
        a = de.OdeSystem(de.rhs_plugins.nbody, y, t=(0.0, 0.1), dt=0.01, constants=consts, ensemble=True)
        a.method = "BABS9O7H"
        a.integrate()
        return a
    a = run(y0)
    y1 = a[-1].y
    assert bool((a.accepted == 10).all()) and bool((a.status == 0).all())
    # total momentum of every system is conserved to roundoff (pairwise forces cancel)
    md = torch.as_tensor(m).cuda()[:, None, None]
    assert float(((md * y1[1]).sum(0) - (md * y0[1]).sum(0)).abs().max()) < 2e-14
    # energy drift of the 7th-order scheme over 10 steps, sampled systems
    def energy(y, cols):
        q, v = y[0][:, :, cols].double(), y[1][:, :, cols].double()
        kin = 0.5 * (md * v * v).sum((0, 1))
        d = q[:, None] - q[None, :]
        r = torch.sqrt((d * d).sum(2) + 1e-4)
        pot = -0.5 * ((md[:, None, 0] * md[None, :, 0]) / r * (1 - torch.eye(nb, device="cuda")[:, :, None])).sum((0, 1))
        return kin + pot
    cols = torch.arange(0, B, 257, device="cuda")
    e0, e1 = energy(y0, cols), energy(y1, cols)
    drift = ((e1 - e0) / e0.abs()).abs()
    # roundoff for almost every system; a close encounter (softening 1e-4, dt = 0.01) costs a few 1e-4 in the worst one
    assert float(drift.median()) < 1e-12 and float(drift.max()) < 1e-2
    # a subset of the systems on its own: same bits; and two runs are bitwise equal
    sub = slice(3, B, 1021)
    b = run(y0[..., sub].contiguous())
    assert torch.equal(b[-1].y, y1[..., sub])
    c = run(y0)
    assert torch.equal(c[-1].y, y1)


def test_config2_linear_4096_by_8192():
    """y' = A y, 4096 x 4096, 8192 columns, Dormand-Prince 8(7), A @ Y_stage on the library's own tensor-core GEMM:
    a block of columns integrated on its own gives the same bits as inside the full ensemble (per-column control,
    column-independent GEMM arithmetic, transparent compaction)."""
    de = _de()
    n, cols = 4096, 8192
    A, Y0 = de.benchmarks.linear_system(n, cols)
    Ad, Yd = torch.as_tensor(A).cuda(), torch.as_tensor(Y0).cuda()

    def run(y):
        a = de.OdeSystem(de.rhs_plugins.linear, y, t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, constants=dict(A=Ad), ensemble=True)
        a.method = "RK87"
        a.integrate()
        return a
    de.rhs_plugins.linear.gemm = "dmma"                 # this test is about the FP64 DMMA kernel (the default "auto" picks INT8 here)
    try:
        a = run(Yd)
        assert a.success and bool((a[-1].t == 1.0).all())
        att = (a.accepted + a.rejected)
        assert int(att.min()) >= 8 and int(att.max()) <= 12
        sub = slice(1024, 1024 + 256)
        b = run(Yd[:, sub].contiguous())
    finally:
        de.rhs_plugins.linear.gemm = "auto"
    assert torch.equal(b.accepted, a.accepted[sub]) and torch.equal(b.rejected, a.rejected[sub])
    assert float((b[-1].y - a[-1].y[:, sub]).abs().max()) <= 1e-13 * float(a[-1].y.abs().max())
    # the flow e^{A t} of the damped near-skew matrix contracts: |y(1)| <= e^{-0.1} |y(0)| up to the integration error
    ratio = a[-1].y.norm(dim=0) / Yd.norm(dim=0)
    assert float((ratio - np.exp(-0.1)).abs().max()) < 1e-8
    # sampled columns against the CPU oracle (each column = one trajectory of the reference semantics), at full size
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    idx = np.linspace(0, cols - 1, 6).astype(int)
    ref = orc.erk_ensemble(orc.RHS_LINEAR, Y0[:, idx], 0.0, 1.0, 1e-2, 1e-9, 1e-9, *tb.rk_arrays("RK8713MSolver"), shared=A,
                           threads=os.cpu_count())
    got = a[-1].y[:, idx].cpu().numpy()
    assert rel_inf(got, ref["y"]).max() <= 1e-10
    assert np.array_equal(a.accepted[idx].cpu().numpy(), ref["accepted"]) and np.array_equal(a.rejected[idx].cpu().numpy(), ref["rejected"])
    # no product of the configuration left the library's own tensor-core kernel (compacted tail included)
    assert de.rhs_plugins.linear.calls_dsb > 0
    lib_before = de.rhs_plugins.linear.calls_library
    de.rhs_plugins.linear.gemm = "dmma"
    try:
        a.reset()
        a.integrate()
    finally:
        de.rhs_plugins.linear.gemm = "auto"
    assert de.rhs_plugins.linear.calls_library == lib_before


def test_config2_linear_on_the_int8_tensor_cores():
    """Config 2 with `linear.gemm = "int8"` (csrc/ozaki_gemm.cu, CTA pairs): the digit planes of a column depend on that
    column alone and the accumulators are exact, so a block of columns integrated on its own gives the same BITS as inside
    the full ensemble; sampled columns against the CPU oracle (counts identical, states within the parity bar); every
    product of the run on the INT8 kernel or (compacted tail) on dsb_dgemm, none on the library."""
    de = _de()
    lin = de.rhs_plugins.linear
    n, cols = 4096, 8192
    A, Y0 = de.benchmarks.linear_system(n, cols)
    Ad, Yd = torch.as_tensor(A).cuda(), torch.as_tensor(Y0).cuda()

    def run(y):
        a = de.OdeSystem(lin, y, t=(0.0, 1.0), dt=1e-2, rtol=1e-9, atol=1e-9, constants=dict(A=Ad), ensemble=True)
        a.method = "RK87"
        a.integrate()
        return a
    lin.gemm = "int8"
    try:
        int8_before, lib_before = lin.calls_int8, lin.calls_library
        a = run(Yd)
        assert lin.calls_int8 > int8_before and lin.calls_library == lib_before
        assert a.success and bool((a[-1].t == 1.0).all())
        sub = slice(2048, 2048 + 256)
        b = run(Yd[:, sub].contiguous())
    finally:
        lin.gemm = "auto"
        lin._int8.clear()
    assert torch.equal(b.accepted, a.accepted[sub]) and torch.equal(b.rejected, a.rejected[sub])
    assert torch.equal(b[-1].y, a[-1].y[:, sub])
    ratio = a[-1].y.norm(dim=0) / Yd.norm(dim=0)
    assert float((ratio - np.exp(-0.1)).abs().max()) < 1e-8
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    idx = np.linspace(0, cols - 1, 6).astype(int)
    ref = orc.erk_ensemble(orc.RHS_LINEAR, Y0[:, idx], 0.0, 1.0, 1e-2, 1e-9, 1e-9, *tb.rk_arrays("RK8713MSolver"), shared=A,
                           threads=os.cpu_count())
    assert rel_inf(a[-1].y[:, idx].cpu().numpy(), ref["y"]).max() <= 1e-10
    assert np.array_equal(a.accepted[idx].cpu().numpy(), ref["accepted"]) and np.array_equal(a.rejected[idx].cpu().numpy(), ref["rejected"])


def test_config1_events_are_pure_observers_at_full_size():
    """2^20 Lorenz trajectories with two recorded sections: integration bit-identical to the run without events, every
    recorded state on its hyperplane, event times increasing per trajectory."""
    de = _de()
    n = 1 << 20
    y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(n)).cuda()
    cc = de.events.component_crossing
    evs = [cc(0, 0.0, dim=3), cc(1, 0.0, dim=3, direction=-1)]

    def run(events):
        a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
        a.method = "RK45"
        a.event_capacity = 16
        a.integrate(events=events)
        return a
    a, b = run(None), run(evs)
    assert torch.equal(a[-1].y, b[-1].y) and torch.equal(a.accepted, b.accepted) and torch.equal(a.rejected, b.rejected)
    log = b.events
    assert not log.overflowed and int(log.count.sum()) > n
    k = torch.arange(log.capacity, device="cuda")[None, :]
    valid = k < log.count[:, None]
    g = torch.where(log.index == 0, log.y[0], log.y[1])           # both events are bare components
    assert float(g[valid].abs().max()) <= 1e-9
    t = log.t
    later = valid[:, 1:]
    assert bool((t[:, 1:][later] >= t[:, :-1][later]).all())
