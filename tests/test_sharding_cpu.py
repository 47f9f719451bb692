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


def test_bench_bank_tables_and_partitions():
    """bench.py's bank description and its strong-scaling partition (host logic only): the mixed bank of configs[4] is
    sorted by (samplesPerBaud, constelationSize), its cost-balanced ranges tile it exactly once with near-equal cost, and
    the algorithmic bytes follow SURVEY.md 8d."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    w5, w0 = bench.WORKLOADS["config5"], bench.WORKLOADS["bank8psk"]
    t5 = bench.channel_table(w5)
    assert len(t5) == 8192
    keys = [(p["samplesPerBaud"], p["constelationSize"]) for p in t5]
    assert keys == sorted(keys) and {k[0] for k in keys} == {8, 9, 10} and {k[1] for k in keys} == {2, 4, 8}
    assert {p["numAvg"] for p in t5} == {50, 100, 200} and {p["phaseAvg"] for p in t5} == {25, 50, 100}
    assert t5 == bench.channel_table(w5)                      # seeded: every rank builds the same table
    costs = [bench.channel_cost(p, w5["samples"]) for p in t5]
    for world in (2, 4, 8):
        r = balanced_ranges(costs, world)
        assert r[0][0] == 0 and r[-1][1] == 8192 and all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        sums = [sum(costs[lo:hi]) for lo, hi in r]
        assert max(sums) <= 1.02 * sum(costs) / world
    t0 = bench.channel_table(w0)
    assert len(t0) == 4096 and all(p == t0[0] for p in t0)
    # 8 N + K (8 + 4 + 2 + 2 b) per channel, K = N / S in steady state
    assert bench.algorithmic_bytes(t0[:1], 1_000_000) == 8_000_000 + 125_000 * (14 + 6)
    assert bench.algorithmic_bytes(t0, 1_000_000) == 43_008_000_000
    assert bench.sample_rows(4096, 512)[:3] == [0, 8, 16] and len(set(bench.sample_rows(8192, 512))) == 512
