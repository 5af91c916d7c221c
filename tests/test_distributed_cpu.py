"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes shard an ensemble round-robin,
all-reduce the termination statistics and gather it back (desolver_b200/distributed.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, n_global, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from desolver_b200 import distributed as dd
    g = torch.arange(3 * n_global, dtype=torch.float64).reshape(3, n_global)
    shard = dd.shard_round_robin(g)
    assert shard.shape[-1] == dd.local_count(n_global, rank, world_size)
    assert torch.equal(shard, g[:, rank::world_size])
    # pretend every trajectory took (index % 7) + 1 accepted steps and index % 3 rejections; two failed on rank 1
    idx = torch.arange(n_global)[rank::world_size]
    acc = (idx % 7 + 1).to(torch.int32)
    rej = (idx % 3).to(torch.int32)
    status = torch.zeros_like(acc)
    if rank == 1:
        status[:2] = 1
    stats = dd.reduce_statistics(acc, rej, status)
    full_idx = torch.arange(n_global)
    assert stats["accepted"] == int((full_idx % 7 + 1).sum())
    assert stats["attempts"] == int((full_idx % 7 + 1).sum() + (full_idx % 3).sum())
    assert stats["failed"] == 2 and stats["unfinished"] == 0
    back = dd.gather_ensemble(shard * 2.0, n_global)
    assert torch.equal(back, g * 2.0)
    cnt = dd.gather_ensemble(acc, n_global)
    assert torch.equal(cnt, (full_idx % 7 + 1).to(torch.int32))
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("n_global", [10, 11])
def test_shard_reduce_gather_gloo_world2(n_global):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_global, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() == "ok"


def test_single_process_fallbacks():
    from desolver_b200 import distributed as dd
    x = torch.arange(12.0).reshape(2, 6)
    assert torch.equal(dd.shard_round_robin(x, 1, 3), x[:, 1::3])
    assert dd.local_count(10, 0, 3) == 4 and dd.local_count(10, 2, 3) == 3
    assert torch.equal(dd.gather_ensemble(x, 6), x)
