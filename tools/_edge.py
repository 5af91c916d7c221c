import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import desolver_b200 as de
cc = de.events.component_crossing
# 1/2: host_io with pageable memory and tiny / ragged sizes, both pipelines
for n in (1, 31, 33, 1000):
    y0 = de.benchmarks.lorenz_ensemble(n, seed=3)
    ref = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0).cuda(), t=(0, 0.5), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
    ref.integrate()
    for mode in ("streamed", "chunked"):
        h = de.OdeSystem(de.rhs_plugins.lorenz, torch.as_tensor(y0), t=(0, 0.5), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True, host_io=True)
        h.host_pipeline = mode
        h.integrate()
        assert torch.equal(h[-1].y, ref[-1].y.cpu()) and torch.equal(h.accepted, ref.accepted.cpu()), (n, mode)
print("host_io edge sizes ok")
# 3: event store overflow
y0 = torch.as_tensor(de.benchmarks.lorenz_ensemble(500, seed=4)).cuda()
a = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0, 5.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
a.event_capacity = 2
a.integrate(events=[cc(0, 0.0, dim=3), cc(1, 0.0, dim=3)])
log = a.events
assert log.overflowed and int(log.count.max()) > 2 and len(log.for_trajectory(int(log.count.argmax()))) == 2
print("overflow ok: max count", int(log.count.max()))
# 4: strict flavour with events, fused and stage path agree on the event list
res = []
for fused in (True, False):
    b = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True, strict=True)
    b.use_fused = fused
    b.integrate(events=[cc(0, 0.0, dim=3)])
    res.append((b.events.count.clone(), b.events.t.clone(), b[-1].y.clone()))
assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][2], res[1][2])
m = torch.arange(res[0][1].shape[1], device="cuda")[None, :] < res[0][0][:, None]
print("strict fused vs stage events: counts equal, max |dt| %.3g" % float((res[0][1] - res[1][1])[m].abs().max()))
# 5: single system, terminal event in the very first step
s = de.OdeSystem(de.rhs_plugins.harmonic_oscillator, np.array([1.0, 0.0]), t=(0, 1.0), dt=0.5, rtol=1e-9, atol=1e-9)
ev = cc(0, 0.99999, dim=2, direction=-1, is_terminal=True)
s.integrate(events=ev)
print("first-step terminal:", s.integration_status, float(s[-1].t), [float(x.t) for x in s.events])
# 6: continue after an event-free first leg, then events on the second leg
c = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
c.integrate(1.0)
c.integrate(2.0, events=[cc(0, 0.0, dim=3)])
d = de.OdeSystem(de.rhs_plugins.lorenz, y0, t=(0, 2.0), dt=1e-2, rtol=1e-9, atol=1e-9, ensemble=True)
d.integrate(1.0); d.integrate(2.0)
assert torch.equal(c[-1].y, d[-1].y) and float(c.events.t[c.events.count > 0][:, 0].min()) >= 1.0
print("continuation with events ok")
