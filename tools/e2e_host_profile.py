"""Host-side cost of one end-to-end pass (OdeSystem(host_io=True) construction + integrate()): host clock and cProfile.

    python tools/e2e_host_profile.py
"""
import cProfile, pstats, io, time, sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import desolver_b200 as de
from desolver_b200 import benchmarks as bm
cfg = bm.LORENZ
n = 1 << 20
y0 = torch.as_tensor(bm.lorenz_ensemble(n)).pin_memory()
host_out = dict(y=torch.empty((3, n), dtype=torch.float64).pin_memory())
def one():
    s = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"],
                     ensemble=True, device="cuda:0", host_io=True, host_out=host_out, host_chunks=16)
    s.integrate()
    return s
for _ in range(5): one()
torch.cuda.synchronize()
# host clock: constructor / integrate
tc = ti = 0.0
for _ in range(30):
    t0 = time.perf_counter()
    s = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"],
                     ensemble=True, device="cuda:0", host_io=True, host_out=host_out, host_chunks=16)
    t1 = time.perf_counter()
    s.integrate()
    t2 = time.perf_counter()
    tc += t1 - t0; ti += t2 - t1
    s = None
print("constructor %.3f ms  integrate %.3f ms" % (tc / 30 * 1e3, ti / 30 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(50): one()
pr.disable()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(28)
print(st.getvalue()[:6000])
