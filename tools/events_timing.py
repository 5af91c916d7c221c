#!/usr/bin/env python
"""Cost of device event detection on the headline ensemble: 2^20 Lorenz trajectories, RK45CK, t in [0, 2], with
(a) no events, (b) two recorded sections (x = 0 both ways, y = 0 falling), (c) a terminal event (first rising x = 0),
(d) the successive maxima of z (an event on dState/dt)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import desolver_b200 as de  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
cfg = de.benchmarks.LORENZ
y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(n)).cuda()
cc = de.events.component_crossing
cases = {"none": None,
         "two sections": [cc(0, 0.0, dim=3), cc(1, 0.0, dim=3, direction=-1)],
         "terminal x=0 rising + y=0 section": [cc(0, 0.0, dim=3, direction=1, is_terminal=True), cc(1, 0.0, dim=3)],
         "Lorenz map: maxima of z (requires_dstate)": [de.events.component_extremum(2, dim=3, direction=-1)]}
for name, evs in cases.items():
    a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0.0, cfg["tf"]), dt=cfg["dt0"], rtol=1e-9, atol=1e-9, ensemble=True)
    a.event_capacity = 16
    best = 1e30
    for rep in range(4):
        a.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a._kernel_events = (e0, e1)
        a.integrate(events=evs)
        torch.cuda.synchronize()
        if rep:
            best = min(best, e0.elapsed_time(e1))
    st = a.stats
    out = dict(case=name, n=n, kernel_ms=best, attempts=st["attempts"], trajectory_steps_per_s=st["attempts"] / best * 1e3)
    if evs:
        log = a.events
        out.update(events_found=int(log.count.sum()), max_per_trajectory=int(log.count.max()), overflowed=log.overflowed,
                   terminated=int((a.status == 3).sum()))
    print(json.dumps(out))
