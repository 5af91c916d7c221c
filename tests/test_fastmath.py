"""CPU check of the controller's fast double-precision building blocks (csrc/fastmath.cuh): the same
header compiles for the host, where the algorithms are compared with libm on 2e6 random inputs."""
import math
import os
import subprocess

from conftest import ROOT


def test_fastmath_accuracy(tmp_path):
    exe = str(tmp_path / "fm")
    subprocess.run(["g++", "-O2", "-o", exe, os.path.join(ROOT, "tests", "fastmath_host.cpp")], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    w10, w16, wrt, wat, wdiv, wmax, wext, w3 = map(float, out[:8])
    eps = 2.220446049250313e-16
    assert w10 <= 2 * eps and w16 <= 2 * eps and wrt <= 2 * eps      # x^(-1/P): <= 2 ulp
    assert wat <= 3 * eps                                            # 1 + atan(u): <= 3 ulp
    assert wdiv <= 2 * eps                                           # division: <= 2 ulp
    assert wmax == 0.0                                               # integer abs-max is exact
    assert wext <= 4 * eps
    assert w3 <= 1e-13          # rare >=3rd-attempt branch: 11th/27th powers of integer roots amplify rounding
    assert float(out[8]) == 1.0 + math.atan(-1.0)
    assert abs(float(out[9]) - (1.0 + math.atan(0.8e300))) <= 2 * eps * 2.57
