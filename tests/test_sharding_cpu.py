"""Host-side logic of the N>1 path on CPU: channel-range sharding (no data-path collective) and
the max-over-ranks timing reduction bench.py uses, exercised with a world_size-2 gloo group."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from psk_soft_b200.shard import balanced_ranges, channel_ranges


def test_channel_ranges_tile_the_bank():
    for n, w in ((4096, 1), (4096, 8), (8192, 8), (10, 4), (3, 8)):
        r = channel_ranges(n, w)
        assert len(r) == w and r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        sizes = [hi - lo for lo, hi in r]
        assert max(sizes) - min(sizes) <= 1


def test_balanced_ranges_even_out_cost():
    costs = [1.0] * 100 + [4.0] * 100          # e.g. 1M-sample and 4M-sample channels
    r = balanced_ranges(costs, 4)
    assert r[0][0] == 0 and r[-1][1] == 200 and all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
    sums = [sum(costs[lo:hi]) for lo, hi in r]
    assert max(sums) <= 1.1 * sum(costs) / 4


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_channels = 4099
    lo, hi = channel_ranges(n_channels, world)[rank]
    # every rank reports its range and a fake step time; rank 0 checks coverage and takes the max
    mine = torch.tensor([lo, hi], dtype=torch.int64)
    allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allr, mine)
    t = torch.tensor([10.0 + 3.0 * rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    units = torch.tensor([float(hi - lo)], dtype=torch.float64)
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    dist.barrier()
    if rank == 0:
        q.put(([x.tolist() for x in allr], float(t.item()), float(units.item())))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ranges, tmax, units = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == 4099
    assert tmax == 13.0 and units == 4099.0
