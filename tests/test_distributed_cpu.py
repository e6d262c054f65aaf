"""Host-side multi-rank logic on CPU: world_size-2 gloo (the N>1 path's sharding and the flat
gradient all-reduce), no GPU needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from toad_b200.distributed import FlatGradBucket, aligned_rounds, shard_slides


def test_shard_slides_partition_and_balance():
    rng = np.random.default_rng(7)
    lengths = rng.integers(5000, 80001, size=512).tolist()
    for world in (1, 2, 4, 8):
        shards = [shard_slides(lengths, r, world) for r in range(world)]
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(512))                         # disjoint cover
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths)           # LPT bound
        rr = [shard_slides(lengths, r, world, balance=False) for r in range(world)]
        assert sorted(i for s in rr for i in s) == list(range(512))
        assert all(all(i % world == r for i in s) for r, s in enumerate(rr))


def test_aligned_rounds_cover_and_straggler_bound():
    """Length-bucketed synchronous rounds: every slide exactly once, one per rank per round, and the per-round
    max / mean patch count (what the all-reduce waits for) close to 1 -- the unbucketed order pays ~1.7x at world 8."""
    rng = np.random.default_rng(7)
    lengths = rng.integers(5000, 80001, size=512).tolist()
    for world in (1, 2, 4, 8):
        for seed in (None, 3):
            rounds = aligned_rounds(lengths, world, seed=seed)
            assert all(len(r) == world for r in rounds)
            flat = sorted(i for r in rounds for i in r if i >= 0)
            assert flat == list(range(512))
            cost = sum(max(lengths[i] for i in r if i >= 0) for r in rounds)      # sum of per-step maxima
            ideal = sum(lengths) / world
            assert cost <= 1.03 * ideal, (world, cost / ideal)
    naive = sum(max(lengths[s * 8 + r] for r in range(8)) for s in range(64))
    assert naive > 1.5 * sum(lengths) / 8                                           # what bucketing removes
    # tail group padded with -1; order differs with the seed but the partition does not
    r5 = aligned_rounds(lengths[:13], 4, seed=1)
    assert len(r5) == 4 and sum(i < 0 for r in r5 for i in r) == 3
    assert aligned_rounds(lengths, 8, seed=1) != aligned_rounds(lengths, 8, seed=2)
    assert aligned_rounds(lengths, 8, seed=1) == aligned_rounds(lengths, 8, seed=1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank)})
    from toad_b200.distributed import init_from_env
    info = init_from_env("gloo")
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(8, 4), nn.ReLU(), nn.Linear(4, 3))   # same weights on every rank
    bucket = FlatGradBucket(model)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    x = torch.full((5, 8), float(rank + 1))
    loss = model(x).pow(2).sum()
    loss.backward()
    local = bucket.flat.clone()
    bucket.allreduce(average=True)
    # every p.grad is still a view of the flat buffer after backward + all-reduce
    off = 0
    ok_alias = True
    for p in bucket.params:
        ok_alias &= p.grad.data_ptr() == bucket.flat[off:off + p.numel()].data_ptr()
        off += p.numel()
    opt.step()
    bucket.zero()
    after = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    q.put((rank, info["world"], local.numpy(), bucket.flat.abs().sum().item(), ok_alias, after.numpy()))
    # reduced gradient must equal the mean of the per-rank local gradients
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    q.put((rank, "mean", torch.stack(gathered).mean(0).numpy()))
    dist.destroy_process_group()


def test_flat_grad_allreduce_world2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2 * world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    first = {r[0]: r for r in got if r[1] != "mean"}
    assert set(first) == {0, 1}
    assert all(first[r][1] == world for r in first)
    assert all(first[r][4] for r in first)                       # aliasing held
    assert all(first[r][3] == 0.0 for r in first)                # zero() cleared the bucket
    # identical parameters on both ranks after the step (same averaged gradient)
    np.testing.assert_array_equal(first[0][5], first[1][5])
    assert not np.array_equal(first[0][2], first[1][2])          # local grads differed
