"""world_size-2 gloo tests (CPU) of the multi-rank plumbing bench.py uses for --gpus N:
robot sharding, max-over-ranks timing reduction and the final result gather."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_robots = 37
        lo, hi = bench.shard_robots(n_robots, world, rank)
        # each rank "plans" for its robots: result = deterministic function of the robot id
        local = [(True, float(r) * 0.5, r * 3, r + 1) for r in range(lo, hi)]
        bench.barrier(dist, 0)
        t = bench.max_over_ranks(dist, 1.0 + rank * 2.5, 0)
        allres = bench.gather_results(dist, local, world, rank)
        q.put((rank, lo, hi, t, allres))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    (r0, lo0, hi0, t0, all0), (r1, lo1, hi1, t1, all1) = got
    assert (lo0, hi1) == (0, 37) and hi0 == lo1  # disjoint cover
    assert t0 == t1 == 3.5                       # max over ranks
    assert all1 is None and len(all0) == 37
    assert [a[2] for a in all0] == [r * 3 for r in range(37)]  # robot order preserved


def test_shard_robots_partitions():
    sys.path.insert(0, ROOT)
    import bench

    for n in (1, 7, 1024):
        for world in (1, 2, 4, 8):
            parts = [bench.shard_robots(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
