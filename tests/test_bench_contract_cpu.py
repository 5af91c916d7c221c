"""bench.py contract checks that need no GPU: the reference arm (CPU oracle port) prints ONE JSON line with the keys the
driver reads, and its `config` is the dict the GPU arm prints (the driver compares the two)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_line_and_shared_config():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-seconds", "0.3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["value"] > 1e5
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.static_config(bench.N_PER_GPU)       # the dict the GPU arm prints too
    assert line["config"]["workload"] == "lorenz63_rk45ck_2^20_per_gpu"
    ref = line.get("cpu_reference_numpy")
    assert ref is None or ("cpu_a" in ref and "cpu_b" in ref)


def test_gpu_arm_uses_the_same_config_function():
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("static_config(") >= 3                               # definition + both arms


def test_live_reference_parity_bookkeeping():
    """`parity.reference_live`: CPU-A's per-trajectory results (global index, accepted, attempts, final state) against the
    shard a rank holds (trajectory i lives on rank i % N at local index i // N) -- checked here with the CPU oracle standing
    in for the GPU result, and with the reference when baseline/_ref is installed."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from desolver_b200 import benchmarks as bm
    from desolver_b200.integrators import tableaux as tb
    from oracle import oracle as orc
    cfg = bm.LORENZ
    m = 64
    r = orc.erk_ensemble(orc.RHS_LORENZ, bm.lorenz_ensemble(m), cfg["t0"], cfg["tf"], cfg["dt0"], cfg["rtol"], cfg["atol"],
                         *tb.rk_arrays("RK45CKSolver"), params=[cfg["sigma"], cfg["rho"], cfg["beta"]], threads=2)
    picks = [3, 10, 17, 40, 63]
    bench.reference_numpy_numbers.details = [(i, int(r["accepted"][i]), int(r["accepted"][i] + r["rejected"][i]), r["y"][:, i].copy())
                                             for i in picks]
    try:
        one = bench.reference_live_parity(r["y"], r["accepted"], r["rejected"], 1, 0)
        assert one["checked_trajectories"] == 5 and one["counts_identical_frac"] == 1.0 and one["max_rel_err"] == 0.0
        for rank in (0, 1):                                              # two ranks: round-robin shards
            part = bench.reference_live_parity(r["y"][:, rank::2], r["accepted"][rank::2], r["rejected"][rank::2], 2, rank)
            assert part["checked_trajectories"] == sum(1 for i in picks if i % 2 == rank) and part["counts_identical_frac"] == 1.0
        wrong = r["accepted"].copy()
        wrong[10] += 1
        assert bench.reference_live_parity(r["y"], wrong, r["rejected"], 1, 0)["counts_identical_frac"] == 0.8
        bench.reference_numpy_numbers.details = None
        assert bench.reference_live_parity(r["y"], r["accepted"], r["rejected"], 1, 0) is None
    finally:
        bench.reference_numpy_numbers.details = None
    rec = bench.reference_numpy_numbers(live=False)                     # the committed build-container record
    assert rec is None or ("cpu_a" in rec and "build container" in rec["where"])
