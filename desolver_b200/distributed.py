"""Multi-GPU sharding of an ensemble: one process per GPU, collectives behind the C ABI (NCCL over NVLink).

The path shards by trajectory and has NO exchange inside a step: trajectory i
lives on rank `i mod world` (round-robin, so every rank sees the same distribution of step counts when
parameters are sorted or structured), each rank runs the fused kernel on its shard to completion, and
collectives carry only
  * the termination / statistics all-reduce (attempts, accepted, failed, unfinished trajectories), taken
    straight from the kernel's device counters and enqueued on a side stream behind the kernel -- the host
    does not look at anything until somebody asks for the statistics, and
  * the final state / counter gather: either `dsb_comm_gather` (one in-place ncclAllGather + one interleave
    kernel per tensor), or no collective at all -- `scatter_store=True` lets the fused kernel write every
    finished trajectory straight into slot `i * W + rank` of the root's buffer over NVLink (CUDA IPC), so
    the gather overlaps the integration.

`torch.distributed` is the plumbing (rendezvous, broadcasting the ncclUniqueId / IPC handles); CUDA shards use
`libdesolver_b200.so`'s `dsb_comm_*` entry points, CPU tensors (the world_size-2 gloo tests of the host logic)
use torch's collectives.
"""
import torch
import torch.distributed as dist


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_round_robin(x, rank=None, world_size=None, group=None):
    """Columns rank, rank+W, rank+2W, ... of the trailing (ensemble) axis, as a contiguous tensor."""
    if rank is None or world_size is None:
        rank, world_size = world(group)
    return x[..., rank::world_size].contiguous()


def local_count(n_global, rank, world_size):
    return (n_global - rank + world_size - 1) // world_size


def bind_to_gpu_numa_node(device_index):
    """Pins this process (and therefore the pinned host buffers it allocates afterwards: first-touch placement) to the
    CPUs of the NUMA node the GPU hangs off.  With one process per GPU streaming host-resident ensembles through its
    own GPU (OdeSystem(host_io=True)), eight ranks otherwise share whatever node they were started on.  Returns the node
    (or None if the topology cannot be read)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(int(device_index))).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = "/sys/bus/pci/devices/%s/numa_node" % bus[-12:].lower()
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


_comms = {}


def communicator(device, group=None):
    """The `cuda_backend.Comm` (dsb_comm_handle) of this process for `group`: rank 0 draws a ncclUniqueId, the
    process group broadcasts it, every rank joins.  Cached per (group, device)."""
    from .backend import cuda_backend as cb
    device = torch.device(device)
    key = (id(group) if group is not None else None, device.index)
    comm = _comms.get(key)
    if comm is None:
        rank, world_size = world(group)
        box = [cb.Comm.unique_id() if rank == 0 else None]
        if world_size > 1:
            src = dist.get_global_rank(group, 0) if group is not None else 0
            dist.broadcast_object_list(box, src=src, group=group)
        with torch.cuda.device(device):
            comm = _comms[key] = cb.Comm(device.index or 0, world_size, rank, box[0])
    return comm


def destroy_communicators():
    for c in _comms.values():
        c.close()
    _comms.clear()


def reduce_statistics(accepted, rejected, status, group=None):
    """All-reduce (SUM) of [attempts, accepted, failed, unfinished] over the ranks: the global
    termination test.  Returns a dict of python ints; identical on every rank."""
    stats = torch.stack([
        (accepted.sum(dtype=torch.int64) + rejected.sum(dtype=torch.int64)),
        accepted.sum(dtype=torch.int64),
        (status == 1).sum(dtype=torch.int64),
        (status == 2).sum(dtype=torch.int64),
    ])
    _, world_size = world(group)
    if world_size > 1:
        if stats.is_cuda:
            from .backend import cuda_backend as cb
            comm = communicator(stats.device, group)
            comm.allreduce_counters(stats.data_ptr(), stats.data_ptr(), 4, cb.current_stream_ptr(stats.device))
        else:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    s = stats.tolist()
    return dict(attempts=int(s[0]), accepted=int(s[1]), failed=int(s[2]), unfinished=int(s[3]))


def gather_ensemble(shard, n_global, group=None):
    """Inverse of `shard_round_robin`: every rank receives the full (..., n_global) tensor,
    global[..., r::W] = shard_r.  CUDA shards of 4- or 8-byte elements go through `dsb_comm_gather` (one in-place
    ncclAllGather + one interleave kernel); anything else through torch's all_gather."""
    rank, world_size = world(group)
    if world_size == 1:
        return shard
    lead = tuple(shard.shape[:-1])
    if shard.is_cuda and shard.element_size() in (4, 8):
        from .backend import cuda_backend as cb
        comm = communicator(shard.device, group)
        rows = 1
        for s in lead:
            rows *= int(s)
        flat = shard.contiguous().reshape(rows, shard.shape[-1])
        staging = torch.empty(comm.staging_bytes(rows, flat.element_size(), n_global), dtype=torch.uint8, device=shard.device)
        full = torch.empty((rows, n_global), dtype=shard.dtype, device=shard.device)
        comm.gather(flat, staging, full, n_global, cb.current_stream_ptr(shard.device))
        staging.record_stream(torch.cuda.current_stream(shard.device))
        return full.reshape(lead + (n_global,))
    n_max = (n_global + world_size - 1) // world_size
    padded = torch.zeros(lead + (n_max,), dtype=shard.dtype, device=shard.device)
    padded[..., :shard.shape[-1]] = shard
    flat = torch.empty(world_size * padded.numel(), dtype=shard.dtype, device=shard.device)
    dist.all_gather_into_tensor(flat, padded.contiguous().reshape(-1), group=group)
    out = flat.reshape((world_size,) + lead + (n_max,))
    full = torch.empty(lead + (n_global,), dtype=shard.dtype, device=shard.device)
    for r in range(world_size):
        full[..., r::world_size] = out[r][..., :local_count(n_global, r, world_size)]
    return full


class ScatterTarget:
    """Destination of the fused kernel's scatter-store on the ROOT rank: one CUDA-IPC allocation holding
    y[dim][N], t[N], dt[N], accepted[N], rejected[N], status[N] of the GLOBAL ensemble.  Every rank maps it and passes
    the mapped pointers to its launches (dsb_erk_run_args.out_*), which write trajectory i of rank r to slot
    i * W + r -- over NVLink for r != root, while the integration is still running."""

    def __init__(self, dim, n_global, device, group=None, root=0):
        from .backend import cuda_backend as cb
        self.rank, self.world_size = world(group)
        self.dim, self.n_global, self.root = int(dim), int(n_global), int(root)
        self.device = torch.device(device)
        n = self.n_global
        pad = lambda b: (b + 255) // 256 * 256                      # noqa: E731
        self.offsets, off = {}, 0
        for name, size in (("y", 8 * self.dim * n), ("t", 8 * n), ("dt", 8 * n), ("accepted", 4 * n), ("rejected", 4 * n),
                           ("status", 4 * n)):
            self.offsets[name] = off
            off += pad(size)
        self.nbytes = off
        box = [None]
        if self.rank == self.root:
            self.buffer = cb.PeerBuffer.allocate(self.nbytes, self.device.index or 0)
            box[0] = self.buffer.handle
        if self.world_size > 1:
            src = dist.get_global_rank(group, self.root) if group is not None else self.root
            dist.broadcast_object_list(box, src=src, group=group)
        if self.rank != self.root:
            self.buffer = cb.PeerBuffer.open(box[0], self.nbytes, self.device.index or 0)

    def pointer(self, name):
        return self.buffer.ptr + self.offsets[name]

    def result(self):
        """Root only: the gathered tensors (views of the peer buffer)."""
        if self.rank != self.root:
            return None
        n, b = self.n_global, self.buffer
        return dict(y=b.tensor(torch.float64, (self.dim, n), self.offsets["y"]),
                    t=b.tensor(torch.float64, (n,), self.offsets["t"]),
                    dt=b.tensor(torch.float64, (n,), self.offsets["dt"]),
                    accepted=b.tensor(torch.int32, (n,), self.offsets["accepted"]),
                    rejected=b.tensor(torch.int32, (n,), self.offsets["rejected"]),
                    status=b.tensor(torch.int32, (n,), self.offsets["status"]))


class PushTarget:
    """Destination of the chunked peer copies of `OdeSystem._run_push` on the ROOT rank: one CUDA-IPC allocation holding,
    per tensor, the shards in rank-major order -- y[W][dim][n_max], t[W][n_max], accepted / rejected / status[W][n_max] --
    so that every copy is one contiguous stream over NVLink.  `result()` (root) interleaves them into global order with
    one kernel per tensor (dsb_interleave_shards)."""

    FIELDS = (("y", 8), ("t", 8), ("accepted", 4), ("rejected", 4), ("status", 4))

    def __init__(self, dim, n_global, device, group=None, root=0):
        from .backend import cuda_backend as cb
        self.rank, self.world_size = world(group)
        self.dim, self.n_global, self.root = int(dim), int(n_global), int(root)
        self.device = torch.device(device)
        W = self.world_size
        self.n_max = (self.n_global + W - 1) // W
        pad = lambda b: (b + 255) // 256 * 256                      # noqa: E731
        self.offsets, off = {}, 0
        for name, size in self.FIELDS:
            rows = self.dim if name == "y" else 1
            self.offsets[name] = off
            off += pad(W * rows * self.n_max * size)
        self.nbytes = off
        box = [None]
        if self.rank == self.root:
            self.buffer = cb.PeerBuffer.allocate(self.nbytes, self.device.index or 0)
            box[0] = self.buffer.handle
        if W > 1:
            src = dist.get_global_rank(group, self.root) if group is not None else self.root
            dist.broadcast_object_list(box, src=src, group=group)
        if self.rank != self.root:
            self.buffer = cb.PeerBuffer.open(box[0], self.nbytes, self.device.index or 0)

    def slot(self, name):
        """Address of this rank's slot of tensor `name` in the root's staging buffer."""
        rows, size = (self.dim, 8) if name == "y" else (1, dict(self.FIELDS)[name])
        return self.buffer.ptr + self.offsets[name] + self.rank * rows * self.n_max * size

    def result(self):
        """Root only: the global tensors, interleaved from the staged shards on the current stream."""
        from .backend import cuda_backend as cb
        if self.rank != self.root:
            return None
        out = {}
        st = cb.current_stream_ptr(self.device)
        with torch.cuda.device(self.device):
            for name, size in self.FIELDS:
                rows = self.dim if name == "y" else 1
                full = torch.empty((rows, self.n_global) if name == "y" else (self.n_global,),
                                   dtype=torch.float64 if size == 8 else torch.int32, device=self.device)
                cb.check(cb.lib().dsb_interleave_shards(self.buffer.ptr + self.offsets[name], full.data_ptr(), rows, size,
                                                        self.n_global, self.world_size, st), "dsb_interleave_shards")
                out[name] = full
        return out


class ShardedEnsemble:
    """Round-robin shard of a global ensemble integrated by this rank's `OdeSystem`.

    `make_system(y0_shard, constants_shard)` builds the rank-local system; per-trajectory constants
    (tensors whose last axis has the global ensemble length) are sharded the same way as `y0`.

    `integrate()` enqueues the local integration and, behind it on a side stream, the all-reduce of the kernel's own
    counters; with `wait=False` nothing synchronises with the host until `.statistics` is read (a loop of
    reset()/integrate() then runs rank-asynchronously: the ranks meet in the collective on the device, never on the
    host).  `scatter_store=True` (fused right-hand sides, fresh runs): the kernel stores every finished trajectory
    straight into the root's gather buffer (see ScatterTarget); `gather()` on the root is then free.
    """

    def __init__(self, make_system, y0_global, constants=None, group=None, scatter_store=False, root=0, push=False):
        self.group = group
        self.rank, self.world_size = world(group)
        self.n_global = int(y0_global.shape[-1])
        self.root = int(root)
        consts = {}
        for k, v in (constants or {}).items():
            if isinstance(v, torch.Tensor) and v.ndim >= 1 and v.shape[-1] == self.n_global:
                v = shard_round_robin(v, self.rank, self.world_size)
            consts[k] = v
        self.system = make_system(shard_round_robin(y0_global, self.rank, self.world_size), consts)
        self._global = None                  # device tensor [attempts, accepted, failed, unfinished] after the all-reduce
        self._stats = None
        self._err = None
        self._side = None
        self.scatter = None
        self.push = None
        if push:
            # the gather as chunked peer copies that overlap the integration (see PushTarget / OdeSystem._run_push)
            s = self.system
            self.push = PushTarget(s.numel, self.n_global, s.device, group, self.root)
            s._push_target = dict(n_max=self.push.n_max, **{k: self.push.slot(k) for k, _ in PushTarget.FIELDS})
        if scatter_store:
            s = self.system
            self.scatter = ScatterTarget(s.numel, self.n_global, s.device, group, self.root)
            s._scatter_store = dict(index_mul=self.world_size, index_add=self.rank, stride=self.n_global,
                                    **{k: self.scatter.pointer(k) for k in ("y", "t", "dt", "accepted", "rejected", "status")})

    # ------------------------------------------------------------------ integration
    def integrate(self, t=None, wait=True):
        """Local integration, then the global statistics all-reduce.  A tolerance failure anywhere raises on EVERY
        rank (when the statistics are read: here with wait=True, else at `.statistics`), like the single-GPU error
        policy."""
        from . import exception_types as etypes
        from .backend import cuda_backend as cb
        s = self.system
        self._stats, self._err = None, None
        counters = None
        lazy_before = getattr(s, "lazy_status", False)
        try:
            s.lazy_status = True                 # no host read-back inside integrate()
            s.integrate(t)
            counters = s._pending_counters()     # device int64[4] of this call, or None (stage-level / symplectic path)
        except etypes.FailedIntegration as e:    # keep going: the collective below must still be entered
            self._err = e
        finally:
            s.lazy_status = lazy_before
        if counters is None:
            st = reduce_statistics(s.accepted, s.rejected, s.status, self.group)
            self._stats = st
        else:
            dev = s.device
            self._global = torch.empty(4, dtype=torch.int64, device=dev)
            if self.world_size > 1:
                # side stream: the next launch on the compute stream does not wait for the slowest rank's collective
                if self._side is None:
                    self._side = torch.cuda.Stream(device=dev)
                cur = torch.cuda.current_stream(dev)
                snap = counters.clone()
                self._side.wait_stream(cur)
                comm = communicator(dev, self.group)
                comm.allreduce_counters(snap.data_ptr(), self._global.data_ptr(), 4, C_void(self._side.cuda_stream))
                snap.record_stream(self._side)
                self._global.record_stream(self._side)
            else:
                self._global.copy_(counters)
        if wait:
            return self.statistics
        return None

    @property
    def statistics(self):
        from . import exception_types as etypes
        if self._stats is None and self._global is not None:
            if self._side is not None:
                self._side.synchronize()
            g = self._global.tolist()
            self._stats = dict(attempts=int(g[0]), accepted=int(g[1]), failed=int(g[2]), unfinished=int(g[3]))
        if self._stats is not None and (self._err is not None or self._stats["failed"]):
            err, self._err = self._err, None
            new = etypes.FailedIntegration("Failed to integrate system on %d trajectories" % self._stats["failed"])
            new.__cause__ = err.__cause__ if err is not None else etypes.FailedToMeetTolerances("on another rank")
            failed, self._stats["failed"] = self._stats["failed"], 0      # raise once; the numbers stay readable
            self._stats["failed_trajectories"] = failed
            raise new
        return self._stats

    # ------------------------------------------------------------------ results
    def gather(self):
        """Final state (*dim, N), times, counters of the GLOBAL ensemble: on every rank through `dsb_comm_gather`;
        with `scatter_store=True` on the root only (views of the buffer the kernels wrote, None elsewhere) after a
        device barrier (the all-reduce already enqueued behind every rank's kernel)."""
        s = self.system
        if self.push is not None:
            # every rank's all-reduce sits behind its kernel AND its last peer copy: once it has completed here, the staged
            # shards are complete
            _ = self.statistics
            out = self.push.result()
            if out is not None:
                out["y"] = out["y"].reshape(tuple(s.dim) + (self.n_global,))
            return out
        if self.scatter is not None:
            # the all-reduce sits behind every rank's kernel, so once it has completed here every peer's stores have
            # been issued and flushed (kernel boundary): no further barrier
            _ = self.statistics
            out = self.scatter.result()
            if out is not None:
                out["y"] = out["y"].reshape(tuple(s.dim) + (self.n_global,))
            return out
        return dict(y=gather_ensemble(s[-1].y, self.n_global, self.group),
                    t=gather_ensemble(s[-1].t, self.n_global, self.group),
                    accepted=gather_ensemble(s.accepted, self.n_global, self.group),
                    rejected=gather_ensemble(s.rejected, self.n_global, self.group),
                    status=gather_ensemble(s.status, self.n_global, self.group))


def C_void(ptr):
    import ctypes
    return ctypes.c_void_p(ptr)
