"""Host side of the paged node store (`dsb_node_store`, csrc/node_store.cu): dense output and per-step history.

The reference keeps one `CubicHermiteInterp` object per accepted step in a Python list
(/root/reference/desolver/differential_system.py:170-290, 1020-1023) and one row of `a.t` / `a.y` per accepted step
(`:1015-1018`).  Here both are ONE device-resident log of nodes (trajectory, k, t, y, f) written by the step kernels,
plus an index (CSR offsets + entry locations) built on the device when somebody first reads it.  torch only owns
the memory; every reader and writer is a kernel of libdesolver_b200.so.
"""
import ctypes as C

from .backend import cuda_backend as cb


def _torch():
    import torch
    return torch


class NodeStore(object):
    """`width` = 32 for stores written by the fused kernels (pages claimed per warp), `width` = n for stores written
    by `dsb_node_append_rows` (one row-group per lock-step attempt)."""

    FUSED_WIDTH = 32

    def __init__(self, device, dim, n_traj, kind, capacity, page_entries=256, with_f=True):
        torch = _torch()
        if kind not in ("fused", "rows"):
            raise ValueError("kind must be 'fused' or 'rows'")
        self.kind = kind
        self.device, self.dim, self.n_traj = device, int(dim), int(n_traj)
        self.width = self.FUSED_WIDTH if kind == "fused" else self.n_traj
        self.fields = 2 + 2 * self.dim
        self.page_entries = int(page_entries) if kind == "fused" else self.width
        self.with_f = bool(with_f)
        self.capacity = 0
        self.pool = None
        self.page_fill = None
        self.cursor = torch.zeros(4, dtype=torch.int64, device=device)
        self.node_count = torch.zeros(self.n_traj, dtype=torch.int32, device=device)
        self._index = None
        self.grow(capacity)

    # ------------------------------------------------------------------ memory
    def _round(self, entries):
        unit = self.page_entries
        return max(unit, (int(entries) + unit - 1) // unit * unit)

    @property
    def bytes_per_entry(self):
        return 8 * self.fields

    def grow(self, capacity):
        """Pool of at least `capacity` entries; existing nodes are kept (the log is append-only, so growing appends
        groups)."""
        torch = _torch()
        capacity = self._round(capacity)
        if capacity <= self.capacity:
            return
        pool = torch.empty(capacity * self.fields, dtype=torch.float64, device=self.device)
        if self.pool is not None:
            pool[:self.pool.numel()].copy_(self.pool)
        self.pool = pool
        if self.kind == "fused":
            fill = torch.zeros(capacity // self.page_entries, dtype=torch.int32, device=self.device)
            if self.page_fill is not None:
                fill[:self.page_fill.numel()].copy_(self.page_fill)
            self.page_fill = fill
        self.capacity = capacity

    def clear(self):
        self.cursor.zero_()
        self.node_count.zero_()
        if self.page_fill is not None:
            self.page_fill.zero_()
        self._index = None

    def struct(self):
        s = cb.NodeStoreStruct()
        s.pool, s.capacity, s.width, s.page_entries, s.dim = self.pool.data_ptr(), self.capacity, self.width, self.page_entries, self.dim
        s.cursor, s.node_count = self.cursor.data_ptr(), self.node_count.data_ptr()
        s.page_fill = self.page_fill.data_ptr() if self.page_fill is not None else None
        return s

    def invalidate(self):
        self._index = None

    # ------------------------------------------------------------------ bookkeeping (each reads 32 bytes back)
    def usage(self):
        """(entries claimed, nodes dropped because the pool was full)"""
        c = self.cursor.tolist()
        return int(c[0]), int(c[1])

    def snapshot(self):
        """State needed to undo a run that overflowed the pool (cursor, per-trajectory counts)."""
        return self.cursor.clone(), self.node_count.clone()

    def restore(self, snap):
        cursor, count = snap
        claimed_before = int(cursor[0].item())
        self.cursor.copy_(cursor)
        self.node_count.copy_(count)
        if self.page_fill is not None:                      # pages claimed by the undone run are free again
            self.page_fill[claimed_before // self.page_entries:].zero_()
        self._index = None

    # ------------------------------------------------------------------ index + readers
    def index(self):
        """(offsets[n + 1], loc[total], total): CSR index over the log, built on the device on first use after a run."""
        torch = _torch()
        if self._index is None:
            n, dev = self.n_traj, self.device
            with torch.cuda.device(dev):
                st = cb.current_stream_ptr(dev)
                offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
                scratch = torch.empty(max(1, cb.lib().dsb_node_offsets_scratch(n)), dtype=torch.int64, device=dev)
                cb.check(cb.lib().dsb_node_offsets(self.node_count.data_ptr(), n, offsets.data_ptr(), scratch.data_ptr(), st),
                         "dsb_node_offsets")
                claimed, dropped = self.usage()
                if dropped:
                    raise MemoryError("dense-output node store overflowed: %d nodes were dropped (pool of %d entries)"
                                      % (dropped, self.capacity))
                total = int(offsets[n].item())
                loc = torch.empty(max(1, total), dtype=torch.int64, device=dev)
                bad = torch.zeros(1, dtype=torch.int64, device=dev)
                s = self.struct()
                cb.check(cb.lib().dsb_node_locate(C.byref(s), claimed, n, offsets.data_ptr(), loc.data_ptr(), bad.data_ptr(), st),
                         "dsb_node_locate")
                if int(bad.item()):
                    raise RuntimeError("node store index: %d inconsistent entries" % int(bad.item()))
            self._index = (offsets, loc, total)
        return self._index

    def evaluate(self, tq, stride_traj, stride_query, nq, derivative=False):
        """out[q][d][i]: see dsb_node_eval."""
        torch = _torch()
        offsets, loc, _ = self.index()
        out = torch.empty((int(nq), self.dim, self.n_traj), dtype=torch.float64, device=self.device)
        s = self.struct()
        with torch.cuda.device(self.device):
            cb.check(cb.lib().dsb_node_eval(C.byref(s), offsets.data_ptr(), loc.data_ptr(), self.n_traj, tq.data_ptr(),
                                            int(stride_traj), int(stride_query), int(nq), out.data_ptr(), 1 if derivative else 0,
                                            cb.current_stream_ptr(self.device)), "dsb_node_eval")
        return out

    def gather(self, trajectory, k0=0, count=None, with_f=False):
        """(t[count], y[count][dim][, f[count][dim]]) of nodes k0.. of one trajectory."""
        torch = _torch()
        offsets, loc, _ = self.index()
        have = int((offsets[trajectory + 1] - offsets[trajectory]).item())
        count = have - k0 if count is None else min(int(count), have - k0)
        count = max(count, 0)
        t_out = torch.empty(count, dtype=torch.float64, device=self.device)
        y_out = torch.empty((count, self.dim), dtype=torch.float64, device=self.device)
        f_out = torch.empty((count, self.dim), dtype=torch.float64, device=self.device) if with_f else None
        if count:
            s = self.struct()
            with torch.cuda.device(self.device):
                cb.check(cb.lib().dsb_node_gather(C.byref(s), offsets.data_ptr(), loc.data_ptr(), int(trajectory), int(k0), count,
                                                  t_out.data_ptr(), y_out.data_ptr(), f_out.data_ptr() if with_f else None,
                                                  cb.current_stream_ptr(self.device)), "dsb_node_gather")
        return (t_out, y_out, f_out) if with_f else (t_out, y_out)

    def to_rows(self, with_f=True):
        """The same nodes re-laid out as the row-group log of the stage-level writer (width = n, slot = trajectory, group g
        = node g of every trajectory that has one, header -1 elsewhere): a run that began in the fused kernel continues
        on the stage-level kernels (method switch to a scheme without a fused instantiation, per-component tolerances,
        events next to dense output).  Index arithmetic only -- the node values are copied bit for bit."""
        torch = _torch()
        if self.kind != "fused":
            return self
        n, F, dim, dev = self.n_traj, self.fields, self.dim, self.device
        offsets, loc, total = self.index()
        counts = self.node_count.to(torch.int64)
        groups = int(counts.max().item()) if n else 0
        new = NodeStore(dev, dim, n, "rows", capacity=(groups + 2) * n, with_f=with_f)
        if groups:
            g = torch.arange(groups, device=dev, dtype=torch.int64).unsqueeze(1)                  # (G, 1)
            valid = g < counts.unsqueeze(0)                                                        # (G, n)
            pos = (offsets[:n].unsqueeze(0) + g).clamp_(max=max(total - 1, 0))
            e = loc[pos]                                                                           # entry of node g of trajectory i
            src = (e >> 5) * (F * 32) + (e & 31)
            dst = new.pool[:groups * F * n].view(groups, F, n)
            empty = torch.tensor([-1, 0], dtype=torch.int32, device=dev).view(torch.float64)       # header {trajectory -1, k 0}
            for r in range(F if with_f else 2 + dim):
                vals = self.pool[src + r * 32]
                dst[:, r, :] = torch.where(valid, vals, empty if r == 0 else torch.zeros_like(vals))
        new.cursor[0] = groups * n
        new.node_count.copy_(self.node_count)
        return new

    def append_rows(self, t, y, f, flags_ptr, mask):
        """Stage-level writer: one row-group for the trajectories whose flag word has `mask` set (all if flags_ptr is None)."""
        s = self.struct()
        cb.check(cb.lib().dsb_node_append_rows(C.byref(s), self.n_traj, t.data_ptr(), y.data_ptr(),
                                               f.data_ptr() if f is not None else None, flags_ptr, int(mask),
                                               cb.current_stream_ptr(self.device)), "dsb_node_append_rows")
        self._index = None
