"""Loader of libdesolver_b200.so and the ctypes mirror of include/desolver_b200.h.

The library is the product: if it is missing or cannot be loaded, every entry point of the
package that needs it raises -- there is no eager/PyTorch/CPU fallback.  torch is used for
device memory (`tensor.data_ptr()`), streams (`torch.cuda.current_stream().cuda_stream`) and
`torch.distributed`; all arithmetic of the hot path happens in the library's kernels.
"""
import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_PKG, "libdesolver_b200.so")

ABI_VERSION = 2
DSB_OK, DSB_ERR_BAD_ARGUMENT, DSB_ERR_CUDA, DSB_ERR_UNSUPPORTED = 0, 1, 2, 3
TRAJ_DONE, TRAJ_TOL_FAIL, TRAJ_RUNNING, TRAJ_EVENT = 0, 1, 2, 3
MAX_EVENTS = 4
FLAG_STRICT = 1
MAX_PARAMS = 6
FUSED_STAGES = (1, 2, 4, 6, 7, 13, 17)

_dp = C.POINTER(C.c_double)


class ErkRunArgs(C.Structure):
    """dsb_erk_run_args (include/desolver_b200.h).  Pointers are raw device addresses."""
    _fields_ = [
        ("n", C.c_int64),
        ("y", C.c_void_p), ("t", C.c_void_p), ("dt", C.c_void_p),
        ("params", C.c_void_p * MAX_PARAMS),
        ("param_stride", C.c_int64 * MAX_PARAMS),
        ("tf", C.c_double), ("rtol", C.c_double), ("atol", C.c_double),
        ("accepted", C.c_void_p), ("rejected", C.c_void_p), ("status", C.c_void_p),
        ("max_steps", C.c_int64),
        ("first_call", C.c_int32),
        ("nodes", C.c_void_p), ("node_count", C.c_void_p), ("node_capacity", C.c_int32),
        ("counters", C.c_void_p),
        # ABI 2
        ("stride", C.c_int64),
        ("y_init", C.c_void_p),
        ("t_init", C.c_double), ("dt_init", C.c_double),
        ("fresh", C.c_int32),
        ("done_shift", C.c_int32),
        ("avail", C.c_void_p),
        ("done", C.c_void_p),
        ("n_events", C.c_int32), ("ev_capacity", C.c_int32),
        ("ev_coef", C.c_void_p), ("ev_flags", C.c_void_p),
        ("ev_records", C.c_void_p), ("ev_count", C.c_void_p),
        ("ev_coef_dy", C.c_void_p),
    ]


class ErkHostArgs(C.Structure):
    """dsb_erk_host_args (include/desolver_b200.h)."""
    _fields_ = [
        ("run", ErkRunArgs),
        ("host_y0", C.c_void_p), ("host_y", C.c_void_p),
        ("host_accepted", C.c_void_p), ("host_rejected", C.c_void_p), ("host_status", C.c_void_p),
        ("host_stride", C.c_int64),
        ("chunk_shift", C.c_int32),
        ("flags", C.c_void_p),
        ("stream_up", C.c_void_p), ("stream_down", C.c_void_p),
    ]


MAX_STAGES = 40


class ErkStageArgs(C.Structure):
    """dsb_erk_stage_args (include/desolver_b200.h)."""
    _fields_ = [
        ("n", C.c_int64),
        ("dim", C.c_int32), ("chunks", C.c_int32),
        ("y", C.c_void_p), ("t", C.c_void_p), ("dt", C.c_void_p),
        ("h", C.c_void_p), ("h0", C.c_void_p),
        ("K", C.c_void_p * MAX_STAGES),
        ("y_stage", C.c_void_p), ("t_stage", C.c_void_p), ("dY", C.c_void_p), ("scale", C.c_void_p),
        ("hist", C.c_void_p), ("partial", C.c_void_p),
        ("trial", C.c_void_p), ("flags", C.c_void_p),
        ("tf", C.c_double), ("rtol", C.c_double), ("atol", C.c_double),
        ("accepted", C.c_void_p), ("rejected", C.c_void_p), ("status", C.c_void_p),
        ("counters", C.c_void_p),
        ("one_step", C.c_int32),
        ("n_events", C.c_int32), ("ev_capacity", C.c_int32),
        ("ev_coef", C.c_void_p), ("ev_flags", C.c_void_p),
        ("ev_records", C.c_void_p), ("ev_count", C.c_void_p),
        ("ev_state", C.c_void_p), ("tf_per", C.c_void_p),
        ("ev_coef_dy", C.c_void_p),
    ]


class SympRunArgs(C.Structure):
    """dsb_symp_run_args (include/desolver_b200.h)."""
    _fields_ = [
        ("n", C.c_int64),
        ("y", C.c_void_p), ("t", C.c_void_p), ("dt", C.c_void_p),
        ("tf", C.c_double),
        ("steps", C.c_void_p), ("status", C.c_void_p),
        ("max_steps", C.c_int64),
        ("first_call", C.c_int32), ("bodies", C.c_int32),
        ("masses", C.c_void_p),
        ("G", C.c_double), ("eps2", C.c_double),
    ]


class SympStageArgs(C.Structure):
    """dsb_symp_stage_args (include/desolver_b200.h)."""
    _fields_ = [
        ("n", C.c_int64),
        ("dim", C.c_int32),
        ("y", C.c_void_p), ("t", C.c_void_p), ("dt", C.c_void_p),
        ("h", C.c_void_p), ("dS", C.c_void_p), ("y_stage", C.c_void_p), ("t_stage", C.c_void_p),
        ("flags", C.c_void_p), ("steps", C.c_void_p), ("status", C.c_void_p),
        ("tf", C.c_double),
        ("counters", C.c_void_p),
        ("n_events", C.c_int32), ("ev_capacity", C.c_int32),
        ("ev_coef", C.c_void_p), ("ev_flags", C.c_void_p),
        ("ev_records", C.c_void_p), ("ev_count", C.c_void_p),
        ("ev_state", C.c_void_p), ("tf_per", C.c_void_p), ("ev_scratch", C.c_void_p),
        ("ev_coef_dy", C.c_void_p),
    ]


class LibraryError(RuntimeError):
    """A dsb_* call returned a non-zero status."""


_lib = None


def library_path():
    return _SO


def build(force=False, jobs=None):
    """Compile the library in-tree with nvcc for sm_100a (needs no GPU)."""
    cmd = ["make", "-C", os.path.join(_PKG, "csrc"), "-j%d" % (jobs or os.cpu_count() or 4)]
    if force:
        cmd.append("-B")
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return _SO


def is_available():
    return os.path.exists(_SO)


def lib():
    """The loaded library (ctypes.CDLL) with prototypes set.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise ImportError(
            "libdesolver_b200.so not found at %s: build it with `python -c \"import __graft_entry__ as g; "
            "g.build()\"` or `make -C desolver_b200/csrc`.  desolver_b200 has no CPU fallback." % _SO)
    L = C.CDLL(_SO)
    L.dsb_abi_version.restype = C.c_int
    L.dsb_build_info.restype = C.c_char_p
    L.dsb_last_error.restype = C.c_char_p
    L.dsb_erk_create.restype = C.c_int
    L.dsb_erk_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int,
                                 C.c_double, C.c_uint]
    L.dsb_erk_destroy.restype = C.c_int
    L.dsb_erk_destroy.argtypes = [C.c_void_p]
    L.dsb_erk_run.restype = C.c_int
    L.dsb_erk_run.argtypes = [C.c_void_p, C.POINTER(ErkRunArgs), C.c_void_p]
    L.dsb_erk_run_host.restype = C.c_int
    L.dsb_erk_run_host.argtypes = [C.c_void_p, C.POINTER(ErkHostArgs), C.c_void_p]
    L.dsb_dgemm.restype = C.c_int
    L.dsb_dgemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
    L.dsb_erk_run_launches.restype = C.c_int
    L.dsb_erk_run_launches.argtypes = [C.c_void_p]
    L.dsb_dense_eval.restype = C.c_int
    L.dsb_dense_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_void_p]
    L.dsb_dense_append.restype = C.c_int
    L.dsb_dense_append.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    for name in ("dsb_erk_stage_begin", "dsb_erk_stage_state"):
        getattr(L, name).restype = C.c_int
        getattr(L, name).argtypes = [C.c_void_p, C.POINTER(ErkStageArgs), C.c_int, C.c_void_p]
    L.dsb_erk_stage_finish.restype = C.c_int
    L.dsb_erk_stage_finish.argtypes = [C.c_void_p, C.POINTER(ErkStageArgs), C.c_void_p]
    L.dsb_erk_stage_finish_events.restype = C.c_int
    L.dsb_erk_stage_finish_events.argtypes = [C.c_void_p, C.POINTER(ErkStageArgs), C.c_int, C.c_void_p, C.c_void_p]
    L.dsb_erk_stage_launches.restype = C.c_int
    L.dsb_erk_stage_launches.argtypes = [C.c_void_p]
    L.dsb_symp_create.restype = C.c_int
    L.dsb_symp_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_uint]
    L.dsb_symp_destroy.restype = C.c_int
    L.dsb_symp_destroy.argtypes = [C.c_void_p]
    L.dsb_symp_run.restype = C.c_int
    L.dsb_symp_run.argtypes = [C.c_void_p, C.POINTER(SympRunArgs), C.c_void_p]
    L.dsb_symp_stage_begin.restype = C.c_int
    L.dsb_symp_stage_begin.argtypes = [C.c_void_p, C.POINTER(SympStageArgs), C.c_int, C.c_void_p]
    L.dsb_symp_stage_accumulate.restype = C.c_int
    L.dsb_symp_stage_accumulate.argtypes = [C.c_void_p, C.POINTER(SympStageArgs), C.c_int, C.c_void_p, C.c_void_p]
    L.dsb_symp_stage_commit_events.restype = C.c_int
    L.dsb_symp_stage_commit_events.argtypes = [C.c_void_p, C.POINTER(SympStageArgs), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dsb_symp_stage_commit.restype = C.c_int
    L.dsb_symp_stage_commit.argtypes = [C.c_void_p, C.POINTER(SympStageArgs), C.c_void_p]
    if L.dsb_abi_version() != ABI_VERSION:
        raise ImportError("libdesolver_b200.so has ABI version %d, expected %d" % (L.dsb_abi_version(), ABI_VERSION))
    _lib = L
    return L


def check(status, what=""):
    if status != DSB_OK:
        msg = lib().dsb_last_error().decode("utf-8", "replace")
        raise LibraryError("%s failed with status %d: %s" % (what or "libdesolver_b200 call", status, msg))


def dgemm(A, B, out=None):
    """C = A @ B for contiguous float64 CUDA matrices through the library's FP64 tensor-core kernel (dsb_dgemm).
    Returns None when the shape is outside the kernel's tiling (the caller then uses a library GEMM)."""
    import torch
    M, K = A.shape
    K2, N = B.shape
    if (K != K2 or M % 64 or N % 128 or K % 16 or A.dtype != torch.float64 or B.dtype != torch.float64
            or not (A.is_cuda and B.is_cuda and A.is_contiguous() and B.is_contiguous())
            or A.data_ptr() % 16 or B.data_ptr() % 16):
        return None
    if out is None:
        out = torch.empty((M, N), dtype=torch.float64, device=A.device)
    status = lib().dsb_dgemm(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, current_stream_ptr(A.device))
    if status == DSB_ERR_UNSUPPORTED:
        return None
    check(status, "dsb_dgemm")
    return out


def current_stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class ErkHandle:
    """Owns one dsb_erk_handle: (device, right-hand side, method tableau)."""

    def __init__(self, device_index, rhs_id, dim, tableau_intermediate, tableau_final, order, strict=False):
        import numpy as np
        inter = np.ascontiguousarray(tableau_intermediate, dtype=np.float64)
        final = np.ascontiguousarray(np.atleast_2d(tableau_final), dtype=np.float64)
        self._keep = (inter, final)
        self.stages = inter.shape[0]
        self.final_rows = final.shape[0]
        self.dim = int(dim)
        self.rhs_id = int(rhs_id)
        self.device_index = int(device_index)
        self._h = C.c_void_p()
        check(lib().dsb_erk_create(C.byref(self._h), self.device_index, self.rhs_id, self.dim, self.stages,
                                   inter.ctypes.data_as(_dp), final.ctypes.data_as(_dp), self.final_rows,
                                   float(order), FLAG_STRICT if strict else 0), "dsb_erk_create")

    @property
    def has_fused_kernel(self):
        return lib().dsb_erk_run_launches(self._h) > 0

    def run(self, args, stream_ptr):
        check(lib().dsb_erk_run(self._h, C.byref(args), stream_ptr), "dsb_erk_run")

    def run_host(self, args, stream_ptr):
        """dsb_erk_run_host; returns False when the driver lacks stream memory operations (caller falls back)."""
        status = lib().dsb_erk_run_host(self._h, C.byref(args), stream_ptr)
        if status == DSB_ERR_UNSUPPORTED:
            return False
        check(status, "dsb_erk_run_host")
        return True

    def stage_begin(self, args, first_call, stream_ptr):
        check(lib().dsb_erk_stage_begin(self._h, C.byref(args), 1 if first_call else 0, stream_ptr), "dsb_erk_stage_begin")

    def stage_state(self, args, stage, stream_ptr):
        check(lib().dsb_erk_stage_state(self._h, C.byref(args), int(stage), stream_ptr), "dsb_erk_stage_state")

    def stage_finish(self, args, stream_ptr):
        check(lib().dsb_erk_stage_finish(self._h, C.byref(args), stream_ptr), "dsb_erk_stage_finish")

    def stage_finish_events(self, args, phase, f1_ptr, stream_ptr):
        check(lib().dsb_erk_stage_finish_events(self._h, C.byref(args), int(phase), f1_ptr, stream_ptr),
              "dsb_erk_stage_finish_events")

    @property
    def stage_launches(self):
        return lib().dsb_erk_stage_launches(self._h)

    def close(self):
        if self._h:
            lib().dsb_erk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_handle_cache = {}


def erk_handle(device_index, rhs_id, dim, tableau_intermediate, tableau_final, order, strict=False):
    """Process-wide cache: one dsb_erk_handle per (device, rhs, dim, tableau, flavour).  Creating a
    handle costs a cudaMalloc + synchronous copy, so systems share them."""
    import numpy as np
    key = (int(device_index), int(rhs_id), int(dim), np.asarray(tableau_intermediate).tobytes(),
           np.asarray(tableau_final).tobytes(), float(order), bool(strict))
    h = _handle_cache.get(key)
    if h is None:
        h = _handle_cache[key] = ErkHandle(device_index, rhs_id, dim, tableau_intermediate, tableau_final, order,
                                           strict=strict)
    return h


class SympHandle:
    """Owns one dsb_symp_handle: (device, right-hand side, splitting tableau)."""

    def __init__(self, device_index, rhs_id, dim, drift_count, tableau, strict=False):
        import numpy as np
        tab = np.ascontiguousarray(tableau, dtype=np.float64)
        self.stages = tab.shape[0]
        self._h = C.c_void_p()
        check(lib().dsb_symp_create(C.byref(self._h), int(device_index), int(rhs_id), int(dim), int(drift_count), self.stages,
                                    tab.ctypes.data_as(_dp), FLAG_STRICT if strict else 0), "dsb_symp_create")

    def run(self, args, stream_ptr):
        check(lib().dsb_symp_run(self._h, C.byref(args), stream_ptr), "dsb_symp_run")

    def stage_begin(self, args, first_call, stream_ptr):
        check(lib().dsb_symp_stage_begin(self._h, C.byref(args), 1 if first_call else 0, stream_ptr), "dsb_symp_stage_begin")

    def stage_accumulate(self, args, stage, f_ptr, stream_ptr):
        check(lib().dsb_symp_stage_accumulate(self._h, C.byref(args), int(stage), f_ptr, stream_ptr),
              "dsb_symp_stage_accumulate")

    def stage_commit(self, args, stream_ptr):
        check(lib().dsb_symp_stage_commit(self._h, C.byref(args), stream_ptr), "dsb_symp_stage_commit")

    def stage_commit_events(self, args, phase, f0_ptr, f1_ptr, stream_ptr):
        check(lib().dsb_symp_stage_commit_events(self._h, C.byref(args), int(phase), f0_ptr, f1_ptr, stream_ptr),
              "dsb_symp_stage_commit_events")

    def __del__(self):
        try:
            if self._h:
                lib().dsb_symp_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def symp_handle(device_index, rhs_id, dim, drift_count, tableau, strict=False):
    import numpy as np
    key = ("symp", int(device_index), int(rhs_id), int(dim), int(drift_count), np.asarray(tableau).tobytes(), bool(strict))
    h = _handle_cache.get(key)
    if h is None:
        h = _handle_cache[key] = SympHandle(device_index, rhs_id, dim, drift_count, tableau, strict=strict)
    return h
