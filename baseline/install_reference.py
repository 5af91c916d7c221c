#!/usr/bin/env python
"""Installs the UNMODIFIED reference (desolver 5.1.0, pure Python) into baseline/_ref so that it travels to the GPU box
and its numpy backend can be timed there beside the GPU arm (bench.py `cpu_reference_numpy`, tools/reference_numpy_timing.py).

MEASUREMENT INFRASTRUCTURE, build container only: needs the read-only tree /root/reference.  baseline/_ref is git-ignored.

The plain recipe `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref
/root/reference` fails here: the reference's build backend (`hatchling`, pyproject.toml:1-3) is not in the offline
wheelhouse.  The package is pure Python, so the wheel is built from a COPY under /tmp whose pyproject.toml names
setuptools (installed) as the backend instead; the package sources are byte-identical to /root/reference/desolver (checked
below) and `--no-deps` skips the run-time dependency `autoray` (absent as well; the dispatch-only stand-in
oracle/refshim/autoray replaces it at import time, no arithmetic).
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")

PYPROJECT = """[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"

[project]
name = "desolver"
version = "5.1.0"
requires-python = ">=3.10"
dependencies = ['numpy>=2', 'tqdm>=4.47.0', 'scipy>=1.15.0', 'autoray>=0.7.0', 'einops>=0.8.0']

[tool.setuptools.packages.find]
include = ["desolver*"]
"""


def _same_tree(a, b):
    cmp = filecmp.dircmp(a, b, ignore=["__pycache__"])
    if cmp.left_only or cmp.right_only or cmp.diff_files or cmp.funny_files:
        return False
    return all(_same_tree(os.path.join(a, d), os.path.join(b, d)) for d in cmp.common_dirs)


def install(force=False):
    """Returns the installed tree, or None when there is no reference to install from."""
    pkg = os.path.join(DST, "desolver")
    if os.path.isdir(pkg) and not force:
        return DST
    if not os.path.isdir(os.path.join(SRC, "desolver")):
        return None
    with tempfile.TemporaryDirectory(prefix="desolver_ref_") as tmp:
        shutil.copytree(os.path.join(SRC, "desolver"), os.path.join(tmp, "desolver"),
                        ignore=shutil.ignore_patterns("__pycache__"))
        with open(os.path.join(tmp, "pyproject.toml"), "w") as f:
            f.write(PYPROJECT)
        shutil.rmtree(DST, ignore_errors=True)
        subprocess.run([sys.executable, "-m", "pip", "install", "-q", "--no-index", "--no-build-isolation", "--no-deps",
                        "--find-links", "/opt/wheelhouse", "--target", DST, tmp], check=True, cwd=tmp)
    if not _same_tree(os.path.join(SRC, "desolver"), pkg):
        raise RuntimeError("baseline/_ref/desolver differs from /root/reference/desolver")
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
