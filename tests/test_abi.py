"""The C-ABI library loads on a machine without a GPU and exports every symbol include/*.h declares
(no compute call is made here)."""
import ctypes
import glob
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    names = []
    for hdr in glob.glob(os.path.join(ROOT, "include", "*.h")):
        names += re.findall(r"DSB_API[^;(]*?\b(dsb_\w+)\s*\(", open(hdr).read())
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from desolver_b200.backend import cuda_backend as cb
    if not cb.is_available():
        cb.build()
    lib = ctypes.CDLL(cb.library_path())
    syms = _declared_symbols()
    assert len(syms) >= 8
    for name in syms:
        assert hasattr(lib, name), "missing export %s" % name
    info = cb.lib().dsb_build_info().decode()
    assert "sm_100a" in info and cb.lib().dsb_abi_version() == cb.ABI_VERSION == 3


def test_ctypes_struct_matches_header_layout():
    """dsb_erk_run_args: the ctypes mirror must have the C layout (field order and 8-byte alignment)."""
    from desolver_b200.backend import cuda_backend as cb
    a = cb.ErkRunArgs
    assert a.n.offset == 0 and a.y.offset == 8 and a.params.offset == 32
    assert a.param_stride.offset == 32 + 8 * cb.MAX_PARAMS
    assert a.tf.offset == 32 + 16 * cb.MAX_PARAMS
    assert a.stride.offset == a.counters.offset + 8


def test_ctypes_structs_match_the_compiled_header(tmp_path):
    """Every ctypes mirror has the size and field offsets gcc gives the struct in include/desolver_b200.h."""
    import subprocess
    from desolver_b200.backend import cuda_backend as cb
    pairs = {"dsb_erk_run_args": cb.ErkRunArgs, "dsb_erk_stage_args": cb.ErkStageArgs,
             "dsb_symp_run_args": cb.SympRunArgs, "dsb_symp_stage_args": cb.SympStageArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "desolver_b200.h"', 'int main(void){']
    for cname, cls in pairs.items():
        lines.append('printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "layout")
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, str(src)], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split("\n")
    for line in filter(None, out):
        cname, fname, val = line.split()
        cls = pairs[cname]
        if fname == "sizeof":
            assert ctypes.sizeof(cls) == int(val), line
        else:
            assert getattr(cls, fname).offset == int(val), line


def test_bad_arguments_are_reported_without_a_gpu():
    from desolver_b200.backend import cuda_backend as cb
    lib = cb.lib()
    assert lib.dsb_erk_run(None, None, None) == cb.DSB_ERR_BAD_ARGUMENT
    assert b"null" in lib.dsb_last_error()
    h = ctypes.c_void_p()
    assert lib.dsb_erk_create(ctypes.byref(h), 0, 1, 3, 0, None, None, 2, 5.0, 0) == cb.DSB_ERR_BAD_ARGUMENT


def test_new_entry_points_validate_their_arguments_without_a_gpu():
    from desolver_b200.backend import cuda_backend as cb
    lib = cb.lib()
    assert lib.dsb_dgemm(None, None, None, 64, 128, 16, None) == cb.DSB_ERR_BAD_ARGUMENT
    assert lib.dsb_erk_run_host(None, None, None) == cb.DSB_ERR_BAD_ARGUMENT
    assert lib.dsb_erk_stage_finish_events(None, None, 0, None, None) == cb.DSB_ERR_BAD_ARGUMENT
    assert lib.dsb_symp_stage_commit_events(None, None, 0, None, None, None) == cb.DSB_ERR_BAD_ARGUMENT
    assert b"bad argument" in lib.dsb_last_error() or b"null" in lib.dsb_last_error()


def test_package_has_no_cpu_fallback():
    import torch
    import desolver_b200 as de
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        de.OdeSystem(de.rhs_plugins.lorenz, [1.0, 1.0, 20.0])
