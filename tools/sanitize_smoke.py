#!/usr/bin/env python
"""Small run of every kernel family, meant to be executed under `compute-sanitizer --tool memcheck`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402

y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(777)).cuda()          # not a multiple of the warp size
for strict in (False, True):
    for method in ("RK45", "RK87", "Dormand-Prince", "RK4", "RK108"):
        a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0, 0.3), dt=1e-2, rtol=1e-8, atol=1e-8, ensemble=True, strict=strict,
                         dense_output=(method == "RK45"), dense_capacity=64)
        a.method = method
        a.integrate()
        if method == "RK45":
            a.sol(0.15)
            a.sol(torch.linspace(0.0, 0.3, 777, dtype=torch.float64))
y2, mu = de.benchmarks.vdp_ensemble(333)
a = de.OdeSystem(de.rhs_plugins.van_der_pol, torch.as_tensor(y2), t=(0, 1.0), dt=1e-2, rtol=1e-8, atol=1e-8,
                 constants=dict(mu=torch.as_tensor(mu)), ensemble=True)
a.integrate()
a = de.OdeSystem(de.rhs_plugins.forced_oscillator, torch.as_tensor(y2), t=(0, 1.0), dt=1e-2, rtol=1e-8, atol=1e-8, ensemble=True,
                 dense_output=True, dense_capacity=128)
a.method = "RK1412"
a.integrate()
a.sol(0.5)
A, Y0 = de.benchmarks.linear_system(300, 77)
a = de.OdeSystem(de.rhs_plugins.linear, torch.as_tensor(Y0), t=(0, 0.5), dt=1e-2, rtol=1e-8, atol=1e-8, constants=dict(A=torch.as_tensor(A).cuda()),
                 ensemble=True)
a.method = "RK87"
a.integrate()
state, m = de.benchmarks.nbody_ensemble(37, bodies=24, seed=3)
yn = torch.as_tensor(np.ascontiguousarray(np.moveaxis(state, 1, -1)))
for fused in (True, False):
    a = de.OdeSystem(de.rhs_plugins.nbody, yn, t=(0, 0.035), dt=0.01, constants=dict(m=torch.as_tensor(m), eps2=1e-4, G=1.0), ensemble=True)
    a.method = "BABS9O7H"
    a.use_fused = fused
    a.integrate()
# second-session paths: kernel-side init (reset + rerun), host-resident ensembles (flow-controlled launch and the
# chunked fallback), device event detection (recorded + terminal), time-dependent plugins
a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0, 0.3), dt=1e-2, rtol=1e-8, atol=1e-8, ensemble=True)
a.integrate()
a.reset()
a.integrate()
yh = torch.as_tensor(de.benchmarks.lorenz_ensemble(2500)).pin_memory()
for mode in ("streamed", "chunked"):
    h = de.OdeSystem(de.rhs_plugins.lorenz, yh, t=(0, 0.3), dt=1e-2, rtol=1e-8, atol=1e-8, ensemble=True, host_io=True, host_chunks=3)
    h.host_pipeline = mode
    h.integrate()
    assert h.host_pipeline == mode and bool((h.status == 0).all())
cc = de.events.component_crossing
a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0, 1.0), dt=1e-2, rtol=1e-8, atol=1e-8, ensemble=True)
a.event_capacity = 4
a.integrate(events=[cc(0, 0.0, dim=3), cc(1, 0.0, dim=3, direction=1, is_terminal=True), cc(2, 25.0, dim=3)])
assert int(a.events.count.sum()) > 0
for rhs in (de.rhs_plugins.forced, de.rhs_plugins.duffing):
    a = de.OdeSystem(rhs, torch.as_tensor(y2), t=(0, 1.0), dt=1e-2, rtol=1e-8, atol=1e-8, ensemble=True)
    a.integrate()
a = de.OdeSystem(de.rhs_plugins.harmonic_oscillator, np.array([1.0, 0.0]), t=(0, 1.0), dt=0.05)
a.method = "ABAS5O6H"
a.integrate()
torch.cuda.synchronize()
print("sanitize smoke done")
# the GEMM kernels of the linear right-hand side: DMMA (aligned and edge instantiation) and the INT8 digit-plane product
# (single CTAs for M = 128 / 384, CTA pairs for M = 256)
from desolver_b200.backend import cuda_backend as cb  # noqa: E402
for M, N, K in ((64, 128, 16), (70, 140, 34), (1, 2, 2), (130, 258, 50)):
    Ag = torch.randn(M, K, dtype=torch.float64, device="cuda")
    Bg = torch.randn(K, N, dtype=torch.float64, device="cuda")
    Cg = cb.dgemm(Ag, Bg)
    assert float((Cg - Ag @ Bg).abs().max()) < 1e-12
for M, N, K in ((128, 64, 128), (256, 128, 256), (384, 192, 128)):
    Ag = torch.randn(M, K, dtype=torch.float64, device="cuda")
    Bg = torch.randn(K, N, dtype=torch.float64, device="cuda")
    Cg = cb.Int8Product(Ag)(Bg)
    assert float((Cg - Ag @ Bg).abs().max()) < 1e-12
torch.cuda.synchronize()
print("sanitize_smoke: GEMM kernels done")
