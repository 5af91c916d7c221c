"""Multi-GPU sharding of an ensemble: one process per GPU (`torch.distributed`, NCCL over NVLink).

The path shards by trajectory and has NO exchange inside a step: trajectory i
lives on rank `i mod world` (round-robin, so every rank sees the same distribution of step counts when
parameters are sorted or structured), each rank runs the fused kernel on its shard to completion, and
collectives carry only
  * the termination / statistics all-reduce (attempts, accepted, failed trajectories), and
  * the final state / counter gather.
Nothing here is specific to CUDA tensors, so the host logic is covered by world_size-2 gloo tests on CPU.
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


def reduce_statistics(accepted, rejected, status, group=None):
    """All-reduce (SUM) of [attempts, accepted, failed, unfinished] over the ranks: the global
    termination test.  Returns a dict of python ints; identical on every rank."""
    stats = torch.stack([
        (accepted.sum(dtype=torch.float64) + rejected.sum(dtype=torch.float64)),
        accepted.sum(dtype=torch.float64),
        (status == 1).sum().to(torch.float64),
        (status == 2).sum().to(torch.float64),
    ])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    s = stats.tolist()
    return dict(attempts=int(s[0]), accepted=int(s[1]), failed=int(s[2]), unfinished=int(s[3]))


def gather_ensemble(shard, n_global, group=None):
    """Inverse of `shard_round_robin`: every rank receives the full (..., n_global) tensor.

    Shards are padded to the common length ceil(n_global / W) for one all_gather_into_tensor, then
    interleaved back: global[..., r::W] = shard_r."""
    rank, world_size = world(group)
    if world_size == 1:
        return shard
    n_max = (n_global + world_size - 1) // world_size
    lead = tuple(shard.shape[:-1])
    padded = torch.zeros(lead + (n_max,), dtype=shard.dtype, device=shard.device)
    padded[..., :shard.shape[-1]] = shard
    flat = torch.empty(world_size * padded.numel(), dtype=shard.dtype, device=shard.device)
    dist.all_gather_into_tensor(flat, padded.contiguous().reshape(-1), group=group)
    out = flat.reshape((world_size,) + lead + (n_max,))
    full = torch.empty(lead + (n_global,), dtype=shard.dtype, device=shard.device)
    for r in range(world_size):
        full[..., r::world_size] = out[r][..., :local_count(n_global, r, world_size)]
    return full


class ShardedEnsemble:
    """Round-robin shard of a global ensemble integrated by this rank's `OdeSystem`.

    `make_system(y0_shard, constants_shard)` builds the rank-local system; per-trajectory constants
    (tensors whose last axis has the global ensemble length) are sharded the same way as `y0`.
    """

    def __init__(self, make_system, y0_global, constants=None, group=None):
        self.group = group
        self.rank, self.world_size = world(group)
        self.n_global = int(y0_global.shape[-1])
        consts = {}
        for k, v in (constants or {}).items():
            if isinstance(v, torch.Tensor) and v.ndim >= 1 and v.shape[-1] == self.n_global:
                v = shard_round_robin(v, self.rank, self.world_size)
            consts[k] = v
        self.system = make_system(shard_round_robin(y0_global, self.rank, self.world_size), consts)
        self.statistics = None

    def integrate(self, t=None):
        """Local integration to completion, then the global termination all-reduce.  A tolerance failure
        anywhere raises on EVERY rank (after all ranks finished), like the single-GPU error policy."""
        from . import exception_types as etypes
        err = None
        try:
            self.system.integrate(t)
        except etypes.FailedIntegration as e:      # keep going: the collective below must still be entered
            err = e
        s = self.system
        self.statistics = reduce_statistics(s.accepted, s.rejected, s.status, self.group)
        if err is not None or self.statistics["failed"]:
            new = etypes.FailedIntegration("Failed to integrate system on %d trajectories" % self.statistics["failed"])
            new.__cause__ = err.__cause__ if err is not None else etypes.FailedToMeetTolerances("on another rank")
            raise new
        return self.statistics

    def gather(self):
        """Final state (*dim, N), times, counters of the GLOBAL ensemble on every rank."""
        s = self.system
        return dict(y=gather_ensemble(s[-1].y, self.n_global, self.group),
                    t=gather_ensemble(s[-1].t, self.n_global, self.group),
                    accepted=gather_ensemble(s.accepted, self.n_global, self.group),
                    rejected=gather_ensemble(s.rejected, self.n_global, self.group),
                    status=gather_ensemble(s.status, self.n_global, self.group))
