"""Loader of libdesolver_b200.so and the ctypes mirror of include/desolver_b200.h.

The library is the product: if it is missing or cannot be loaded, every entry point of the
package that needs it raises -- there is no eager/PyTorch/CPU fallback.  torch is used for
device memory (`tensor.data_ptr()`), streams (`torch.cuda.current_stream().cuda_stream`) and
`torch.distributed`; all arithmetic of the hot path happens in the library's kernels.
"""
import ctypes as C
import weakref
import os
import subprocess

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_PKG, "libdesolver_b200.so")

ABI_VERSION = 3
DSB_OK, DSB_ERR_BAD_ARGUMENT, DSB_ERR_CUDA, DSB_ERR_UNSUPPORTED, DSB_ERR_NCCL = 0, 1, 2, 3, 4
COMM_ID_BYTES, IPC_HANDLE_BYTES = 128, 64
TRAJ_DONE, TRAJ_TOL_FAIL, TRAJ_RUNNING, TRAJ_EVENT = 0, 1, 2, 3
MAX_EVENTS = 4            # fused kernels
MAX_STAGE_EVENTS = 16     # stage-level kernels
FLAG_STRICT = 1
MAX_PARAMS = 6
FUSED_STAGES = (1, 2, 4, 6, 7, 13, 17)

_dp = C.POINTER(C.c_double)


class NodeStoreStruct(C.Structure):
    """dsb_node_store (include/desolver_b200.h)."""
    _fields_ = [
        ("pool", C.c_void_p),
        ("capacity", C.c_int64),
        ("width", C.c_int32), ("page_entries", C.c_int32), ("dim", C.c_int32),
        ("cursor", C.c_void_p), ("page_fill", C.c_void_p), ("node_count", C.c_void_p),
    ]


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
        ("store", NodeStoreStruct),
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
        # ABI 3: scatter-store
        ("out_y", C.c_void_p), ("out_t", C.c_void_p), ("out_dt", C.c_void_p),
        ("out_accepted", C.c_void_p), ("out_rejected", C.c_void_p), ("out_status", C.c_void_p),
        ("out_stride", C.c_int64), ("out_index_mul", C.c_int64), ("out_index_add", C.c_int64),
        # ABI 3: stop list
        ("stops", C.c_void_p), ("n_stops", C.c_int32), ("stop_y", C.c_void_p), ("stop_stride", C.c_int64),
    ]


class EventSearchStruct(C.Structure):
    """dsb_event_search (include/desolver_b200.h)."""
    _fields_ = [
        ("n", C.c_int64), ("n_events", C.c_int32),
        ("lo", C.c_void_p), ("hi", C.c_void_p), ("flo", C.c_void_p), ("fhi", C.c_void_p),
        ("x", C.c_void_p), ("tq", C.c_void_p), ("side", C.c_void_p),
        ("t0", C.c_void_p), ("t1", C.c_void_p), ("counter", C.c_void_p),
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
        ("host_t", C.c_void_p),
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
        ("rtol_vec", C.c_void_p), ("atol_vec", C.c_void_p),
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
    L.dsb_ozaki_slices.restype = C.c_int
    L.dsb_ozaki_slices.argtypes = []
    L.dsb_ozaki_split_a.restype = C.c_int
    L.dsb_ozaki_split_a.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dsb_ozaki_split_b.restype = C.c_int
    L.dsb_ozaki_split_b.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dsb_ozaki_gemm.restype = C.c_int
    L.dsb_ozaki_gemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                 C.c_int, C.c_void_p, C.c_void_p]
    L.dsb_erk_run_launches.restype = C.c_int
    L.dsb_erk_run_launches.argtypes = [C.c_void_p]
    ns = C.POINTER(NodeStoreStruct)
    L.dsb_node_offsets_scratch.restype = C.c_int64
    L.dsb_node_offsets_scratch.argtypes = [C.c_int64]
    L.dsb_node_offsets.restype = C.c_int
    L.dsb_node_offsets.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dsb_node_locate.restype = C.c_int
    L.dsb_node_locate.argtypes = [ns, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dsb_node_append_rows.restype = C.c_int
    L.dsb_node_append_rows.argtypes = [ns, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.dsb_node_eval.restype = C.c_int
    L.dsb_node_eval.argtypes = [ns, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                C.c_void_p, C.c_int32, C.c_void_p]
    L.dsb_node_gather.restype = C.c_int
    L.dsb_node_gather.argtypes = [ns, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]
    L.dsb_hermite_eval.restype = C.c_int
    L.dsb_hermite_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]
    for name in ("dsb_erk_stage_begin", "dsb_erk_stage_state"):
        getattr(L, name).restype = C.c_int
        getattr(L, name).argtypes = [C.c_void_p, C.POINTER(ErkStageArgs), C.c_int, C.c_void_p]
    L.dsb_erk_stage_finish.restype = C.c_int
    L.dsb_erk_stage_finish.argtypes = [C.c_void_p, C.POINTER(ErkStageArgs), C.c_void_p]
    L.dsb_erk_stage_finish_events.restype = C.c_int
    L.dsb_erk_stage_finish_events.argtypes = [C.c_void_p, C.POINTER(ErkStageArgs), C.c_int, C.c_void_p, C.c_void_p]
    L.dsb_erk_stage_launches.restype = C.c_int
    L.dsb_erk_stage_launches.argtypes = [C.c_void_p]
    L.dsb_erk_stage_finish_events_located.restype = C.c_int
    L.dsb_erk_stage_finish_events_located.argtypes = [C.c_void_p, C.POINTER(ErkStageArgs), C.c_void_p, C.c_void_p, C.c_void_p]
    es = C.POINTER(EventSearchStruct)
    L.dsb_event_search_begin.restype = C.c_int
    L.dsb_event_search_begin.argtypes = [es, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.dsb_event_search_step.restype = C.c_int
    L.dsb_event_search_step.argtypes = [es, C.c_void_p, C.c_void_p]
    L.dsb_event_search_roots.restype = C.c_int
    L.dsb_event_search_roots.argtypes = [es, C.c_void_p, C.c_void_p]
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
    L.dsb_symp_stage_commit_events_located.restype = C.c_int
    L.dsb_symp_stage_commit_events_located.argtypes = [C.c_void_p, C.POINTER(SympStageArgs), C.c_void_p, C.c_void_p, C.c_void_p,
                                                       C.c_void_p]
    L.dsb_symp_stage_commit.restype = C.c_int
    L.dsb_symp_stage_commit.argtypes = [C.c_void_p, C.POINTER(SympStageArgs), C.c_void_p]
    # collectives and peer buffers (ABI 3)
    L.dsb_comm_available.restype = C.c_int
    L.dsb_comm_nccl_version.restype = C.c_int
    L.dsb_comm_get_unique_id.restype = C.c_int
    L.dsb_comm_get_unique_id.argtypes = [C.c_void_p]
    L.dsb_comm_init_rank.restype = C.c_int
    L.dsb_comm_init_rank.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.dsb_comm_from_nccl.restype = C.c_int
    L.dsb_comm_from_nccl.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int]
    L.dsb_comm_destroy.restype = C.c_int
    L.dsb_comm_destroy.argtypes = [C.c_void_p]
    L.dsb_comm_size.restype = C.c_int
    L.dsb_comm_size.argtypes = [C.c_void_p]
    L.dsb_comm_rank.restype = C.c_int
    L.dsb_comm_rank.argtypes = [C.c_void_p]
    L.dsb_comm_allreduce_counters.restype = C.c_int
    L.dsb_comm_allreduce_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.dsb_comm_gather.restype = C.c_int
    L.dsb_comm_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                  C.c_void_p]
    L.dsb_interleave_shards.restype = C.c_int
    L.dsb_interleave_shards.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p]
    L.dsb_peer_alloc.restype = C.c_int
    L.dsb_peer_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_int64, C.c_void_p, C.c_int]
    L.dsb_peer_free.restype = C.c_int
    L.dsb_peer_free.argtypes = [C.c_void_p, C.c_int]
    L.dsb_peer_open.restype = C.c_int
    L.dsb_peer_open.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int]
    L.dsb_peer_close.restype = C.c_int
    L.dsb_peer_close.argtypes = [C.c_void_p, C.c_int]
    if L.dsb_abi_version() != ABI_VERSION:
        raise ImportError("libdesolver_b200.so has ABI version %d, expected %d" % (L.dsb_abi_version(), ABI_VERSION))
    _lib = L
    return L


def check(status, what=""):
    if status != DSB_OK:
        msg = lib().dsb_last_error().decode("utf-8", "replace")
        raise LibraryError("%s failed with status %d: %s" % (what or "libdesolver_b200 call", status, msg))


def dgemm(A, B, out=None):
    """C = A @ B for contiguous float64 CUDA matrices through the library's FP64 tensor-core kernel (dsb_dgemm): any M,
    even N and K (shapes that are not multiples of its 64 x 128 x 16 tile run the predicated edge instantiation).
    Returns None otherwise (the caller then uses a library GEMM)."""
    import torch
    M, K = A.shape
    K2, N = B.shape
    if (K != K2 or N % 2 or K % 2 or min(M, N, K) < 1 or A.dtype != torch.float64 or B.dtype != torch.float64
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


class Int8Product:
    """C = A @ B for one fixed float64 CUDA matrix A through the INT8 tensor-core path (dsb_ozaki_*, csrc/ozaki_gemm.cu):
    the digit planes of A are cut once (again whenever A is modified in place), those of B per product.  The kernel tiles
    M by 128, K by 128 and N by 64; other shapes are ZERO-PADDED to the tiling in buffers of this object (exact: a zero
    contributes nothing to a row / column maximum nor to a plane product, so the unpadded block of the result is what an
    unpadded kernel would give) at the price of three extra passes over B and C.  `supports(B)` says whether a right-hand
    side can be served at all (K <= 32768: INT32 accumulators); buffers grow to the widest B seen."""

    def __init__(self, A):
        import torch
        if not (A.is_cuda and A.dtype == torch.float64 and A.ndim == 2 and A.is_contiguous()):
            raise ValueError("Int8Product needs a contiguous float64 CUDA matrix")
        self._A = weakref.ref(A)                 # the planes must not keep the matrix alive (rhs_plugins.linear caches us)
        self.M, self.K = A.shape
        self.Mp, self.Kp = -(-self.M // 128) * 128, -(-self.K // 128) * 128
        self.padded = (self.Mp, self.Kp) != (self.M, self.K)
        self.S = lib().dsb_ozaki_slices()
        self.sms = torch.cuda.get_device_properties(A.device).multi_processor_count
        self.a_planes = torch.empty((self.S, self.Mp, self.Kp), dtype=torch.int8, device=A.device)
        self.a_scale = torch.empty(self.Mp, dtype=torch.float64, device=A.device)
        self._a_pad = torch.zeros((self.Mp, self.Kp), dtype=torch.float64, device=A.device) if self.padded else None
        self._a_version = None
        self.b_planes = self.b_scale = self.workspace = self._b_pad = self._c_pad = None
        self.columns = 0

    @property
    def A(self):
        return self._A()

    def shape_ok(self):
        A = self.A
        return A is not None and self.M > 0 and self.K > 0 and self.Kp <= 32768 and A.data_ptr() % 16 == 0

    def supports(self, B):
        import torch
        return (self.shape_ok() and B.ndim == 2 and B.shape[0] == self.K and B.shape[1] > 0
                and B.is_cuda and B.dtype == torch.float64 and B.is_contiguous() and B.device == self.a_planes.device)

    def refresh(self):
        """Cut the planes of A again if it was modified in place since the last split (torch's version counter)."""
        self._prepare(0)

    def _prepare(self, n):
        import torch
        A = self.A
        stream = current_stream_ptr(A.device)
        if self._a_version != A._version:
            src = A
            if self.padded:
                self._a_pad[:self.M, :self.K].copy_(A)
                src = self._a_pad
            check(lib().dsb_ozaki_split_a(src.data_ptr(), self.Mp, self.Kp, self.a_planes.data_ptr(), self.a_scale.data_ptr(),
                                          stream), "dsb_ozaki_split_a")
            self._a_version = A._version
        np_ = -(-n // 64) * 64
        if np_ > self.columns:
            dev = A.device
            self.b_planes = torch.empty((self.S * np_ * self.Kp,), dtype=torch.int8, device=dev)
            self.b_scale = torch.empty(np_, dtype=torch.float64, device=dev)
            self.workspace = torch.empty(np_, dtype=torch.int64, device=dev)
            self.columns = np_
        if n and (self.padded or np_ != n):
            # staging for a right-hand side / result outside the tiling (allocated here, i.e. before any graph capture)
            if self._b_pad is None or self._b_pad.numel() < self.Kp * np_:
                self._b_pad = torch.zeros(self.Kp * self.columns, dtype=torch.float64, device=A.device)
                self._c_pad = torch.empty(self.Mp * self.columns, dtype=torch.float64, device=A.device)

    def __call__(self, B, out=None, dbg_acc=None):
        import torch
        n = B.shape[1]
        np_ = -(-n // 64) * 64
        self._prepare(n)
        staged = self.padded or np_ != n
        if out is None:
            out = torch.empty((self.M, n), dtype=torch.float64, device=B.device)
        stream = current_stream_ptr(B.device)
        L = lib()
        b_src, c_dst = B, out
        if staged:
            b_src = self._b_pad[:self.Kp * np_].view(self.Kp, np_)
            b_src.zero_()
            b_src[:self.K, :n].copy_(B)
            c_dst = self._c_pad[:self.Mp * np_].view(self.Mp, np_)
        check(L.dsb_ozaki_split_b(b_src.data_ptr(), self.Kp, np_, self.b_planes.data_ptr(), self.b_scale.data_ptr(),
                                  self.workspace.data_ptr(), stream), "dsb_ozaki_split_b")
        check(L.dsb_ozaki_gemm(self.a_planes.data_ptr(), self.a_scale.data_ptr(), self.b_planes.data_ptr(), self.b_scale.data_ptr(),
                               c_dst.data_ptr(), self.Mp, np_, self.Kp, self.sms, None if dbg_acc is None else dbg_acc.data_ptr(),
                               stream), "dsb_ozaki_gemm")
        if staged:
            out.copy_(c_dst[:self.M, :n])
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

    def stage_finish_events_located(self, args, f1_ptr, roots_ptr, stream_ptr):
        check(lib().dsb_erk_stage_finish_events_located(self._h, C.byref(args), f1_ptr, roots_ptr, stream_ptr),
              "dsb_erk_stage_finish_events_located")

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

    def stage_commit_events_located(self, args, f0_ptr, f1_ptr, roots_ptr, stream_ptr):
        check(lib().dsb_symp_stage_commit_events_located(self._h, C.byref(args), f0_ptr, f1_ptr, roots_ptr, stream_ptr),
              "dsb_symp_stage_commit_events_located")

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


class Comm:
    """Owns one dsb_comm_handle: an NCCL communicator over the ranks of a sharded ensemble, created from a
    ncclUniqueId that the host side distributes (include/desolver_b200.h, "collectives").  All calls are enqueued on
    the given stream; none synchronises."""

    def __init__(self, device_index, nranks, rank, unique_id):
        if len(unique_id) != COMM_ID_BYTES:
            raise ValueError("unique_id must be %d bytes" % COMM_ID_BYTES)
        self.device_index, self.nranks, self.rank = int(device_index), int(nranks), int(rank)
        self._id = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        self._h = C.c_void_p()
        check(lib().dsb_comm_init_rank(C.byref(self._h), self.device_index, self.nranks, self.rank, self._id),
              "dsb_comm_init_rank")

    @staticmethod
    def available():
        return bool(lib().dsb_comm_available())

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(COMM_ID_BYTES)
        check(lib().dsb_comm_get_unique_id(buf), "dsb_comm_get_unique_id")
        return bytes(buf.raw)

    def allreduce_counters(self, send_ptr, recv_ptr, count, stream_ptr):
        check(lib().dsb_comm_allreduce_counters(self._h, send_ptr, recv_ptr, int(count), stream_ptr),
              "dsb_comm_allreduce_counters")

    def gather(self, shard, staging, full, n_global, stream_ptr):
        """full[r, j * W + w] = shard_w[r, j]: `shard` is this rank's contiguous (rows, n_local) or (n_local,) tensor of
        4- or 8-byte elements, `full` the (rows, n_global) result, `staging` a byte buffer (see staging_bytes)."""
        rows = 1 if shard.ndim == 1 else int(shard.shape[0])
        check(lib().dsb_comm_gather(self._h, shard.data_ptr(), staging.data_ptr(), full.data_ptr(), rows,
                                    shard.element_size(), int(shard.shape[-1]), int(n_global), stream_ptr), "dsb_comm_gather")

    def staging_bytes(self, rows, elem_bytes, n_global):
        return self.nranks * rows * ((n_global + self.nranks - 1) // self.nranks) * elem_bytes

    def close(self):
        if self._h:
            lib().dsb_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PeerBuffer:
    """A device allocation that other processes of the box can map (CUDA IPC): the destination of the fused kernel's
    scatter-store (dsb_erk_run_args.out_*).  The owner calls `PeerBuffer.allocate`, ships `.handle` (64 bytes) to its
    peers, and they call `PeerBuffer.open(handle)`.  `.tensor(dtype, shape, offset)` views the memory as a torch tensor."""

    def __init__(self, ptr, nbytes, device_index, owner, handle=None):
        self.ptr, self.nbytes, self.device_index, self.owner, self.handle = ptr, int(nbytes), int(device_index), owner, handle

    @classmethod
    def allocate(cls, nbytes, device_index):
        ptr = C.c_void_p()
        handle = C.create_string_buffer(IPC_HANDLE_BYTES)
        check(lib().dsb_peer_alloc(C.byref(ptr), int(nbytes), handle, int(device_index)), "dsb_peer_alloc")
        return cls(ptr.value, nbytes, device_index, True, bytes(handle.raw))

    @classmethod
    def open(cls, handle, nbytes, device_index):
        ptr = C.c_void_p()
        buf = C.create_string_buffer(bytes(handle), IPC_HANDLE_BYTES)
        check(lib().dsb_peer_open(C.byref(ptr), buf, int(device_index)), "dsb_peer_open")
        return cls(ptr.value, nbytes, device_index, False)

    def tensor(self, dtype, shape, offset=0):
        """torch view of [offset, offset + prod(shape) * itemsize) of the buffer (through __cuda_array_interface__)."""
        import torch
        typestr = {torch.float64: "<f8", torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1"}[dtype]

        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = dict(shape=tuple(int(x) for x in shape), typestr=typestr,
                                          data=(self.ptr + int(offset), False), version=2)
        t = torch.as_tensor(v, device="cuda:%d" % self.device_index)
        t._dsb_keepalive = self
        return t

    def close(self):
        if self.ptr:
            if self.owner:
                lib().dsb_peer_free(C.c_void_p(self.ptr), self.device_index)
            else:
                lib().dsb_peer_close(C.c_void_p(self.ptr), self.device_index)
            self.ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
