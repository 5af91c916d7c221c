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


def test_caller_defined_rhs_plugins_build_load_and_register_without_a_gpu():
    """rhs_plugins.register_device_rhs: nvcc cross-compiles the plug-in for sm_100a (prebuilt by __graft_entry__.build(),
    found again by its source hash), the plug-in exports the selector, dsb_rhs_register takes it under a user id and
    refuses the built-in id range."""
    import ctypes
    import glob
    import os
    import desolver_b200 as de
    from desolver_b200.backend import cuda_backend as cb
    fn = de.benchmarks.rossler_plugin()
    assert fn is de.benchmarks.rossler and fn.plugin_id >= 1000 and fn.device_plugin in de.rhs_plugins.FUSED
    assert fn.plugin_params == ("a", "b", "c") and fn.plugin_dim == (3,)
    key = fn.device_plugin.split(":")[-1]
    so = glob.glob(os.path.join(os.path.dirname(de.__file__), "_plugins", "dsb_rhs_rossler_%s.so" % key))
    assert len(so) == 1
    plug = ctypes.CDLL(so[0])
    assert hasattr(plug, "dsb_user_rhs_select")
    L = cb.lib()
    L.dsb_rhs_register.restype = ctypes.c_int
    L.dsb_rhs_register.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    sel = ctypes.cast(plug.dsb_user_rhs_select, ctypes.c_void_p)
    assert L.dsb_rhs_register(5, 3, sel) == cb.DSB_ERR_BAD_ARGUMENT and b"dsb_rhs_register" in L.dsb_last_error()
    assert L.dsb_rhs_register(1999, 0, sel) == cb.DSB_ERR_BAD_ARGUMENT
    assert L.dsb_rhs_register(1999, 3, None) == cb.DSB_ERR_BAD_ARGUMENT
    assert L.dsb_rhs_register(1999, 3, sel) == 0
    assert de.benchmarks.rossler_plugin().plugin_id == fn.plugin_id            # registered once, reused
    # the generated source is what the docstring promises: the caller's statements inside an Rhs struct of the library
    cu = open(so[0][:-3] + ".cu").read()
    assert "struct UserRhs" in cu and "dy[2] = p[1] + y[2] * (y[0] - p[2]);" in cu and "select_fused<dsb::UserRhs>" in cu
