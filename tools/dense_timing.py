#!/usr/bin/env python
"""Dense output at ensemble scale: the 2^20-trajectory Lorenz ensemble of the headline with
dense_output=True -- one node (t, y, f) per accepted step in the paged node store -- and the T x N evaluation kernel.
Prints one JSON line: time of the integration with and without the store, bytes written, fraction of the HBM roofline
for the algorithmic 8 * (1 + 2 dim) = 56 B per accepted step, index build and evaluation times."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cfg = de.benchmarks.LORENZ
y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(n)).cuda()


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def system(dense):
    a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(cfg["t0"], cfg["tf"]), dt=cfg["dt0"], rtol=cfg["rtol"], atol=cfg["atol"], ensemble=True,
                     dense_output=dense)
    a.method = "RK45"
    return a


plain = system(False)
plain.integrate()
ms_plain = timed(lambda: (plain.reset(), plain.integrate()))
a = system(True)
a.integrate()
store = a._node_store
claimed, dropped = store.usage()
nodes = int(a.sol.node_count.sum())


def rerun():
    a.reset()
    a.integrate()


ms_dense = timed(rerun)            # includes allocating the pool again (reset() drops the store) and the usage read-back
st = a._node_store
ms_index = timed(lambda: (st.invalidate(), st.index()), reps=2)
grid = torch.linspace(0.05, 1.95, T, dtype=torch.float64, device="cuda")
ms_grid = timed(lambda: a.sol.grid(grid))
ms_one = timed(lambda: a.sol(1.0))
peak = 6538.3
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
extra_ms = ms_dense - ms_plain
print(json.dumps(dict(
    workload="lorenz63_rk45ck_2^20 dense_output=True" if n == 1 << 20 else "lorenz n=%d dense" % n, trajectories=n, nodes=nodes,
    ms_integrate_plain=ms_plain, ms_integrate_dense=ms_dense, store_bytes=int(st.pool.numel() * 8), entries_claimed=int(st.usage()[0]),
    entries_capacity=int(st.capacity), bytes_per_node_written=64, algorithmic_bytes_per_node=56,
    round1_layout_bytes=int(n) * 1024 * 56,
    node_write_GBps_algorithmic=nodes * 56 / ms_dense / 1e6, hbm_frac_algorithmic=nodes * 56 / ms_dense / 1e6 / peak,
    node_write_GBps_over_the_extra_time=(nodes * 56 / extra_ms / 1e6) if extra_ms > 0 else None,
    trajectory_steps_per_s_dense=float(a.stats["attempts"]) / ms_dense * 1e3,
    ms_index_build=ms_index, ms_eval_one_time=ms_one, ms_eval_grid=ms_grid, grid_times=T,
    eval_points_per_s=T * n / ms_grid * 1e3)))
