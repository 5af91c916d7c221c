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
