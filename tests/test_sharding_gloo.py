"""CPU test of the multi-rank host logic with the gloo backend, world_size 2: units are
sharded with no data-path collective; only timing/throughput are reduced."""
import os
import socket
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from libspecbleach_b200.sharding import shard_bounds, shard_units  # noqa: E402


def test_shard_bounds_partition():
    for n in (0, 1, 2, 7, 4096):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                lo, hi = shard_bounds(n, world, r)
                covered += list(range(lo, hi))
                assert 0 <= hi - lo <= n // world + 1
            assert covered == list(range(n))
    assert shard_units(list("abcdefg"), 2, 1) == list("efg")
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from libspecbleach_b200.sharding import job_throughput, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(7, world, rank)             # 7 streams of 60 s each
    audio = 60.0 * (hi - lo)
    ms = 100.0 * (rank + 1)                           # rank 1 is the slow one
    thr, worst = job_throughput(audio, ms, dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, lo, hi, thr, worst))


def test_two_ranks_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 4), (4, 7)]
    for r in res:   # 420 s of audio / max(100, 200) ms
        assert abs(r[3] - 420.0 / 0.2) < 1e-6 and abs(r[4] - 200.0) < 1e-9


def test_time_tiles_cover_the_stream_on_hop_boundaries():
    from libspecbleach_b200.sharding import time_tiles
    n, hop, first = 48000 * 30 + 123, 600, 48000
    for tiles in (1, 2, 3, 8):
        plan = time_tiles(n, hop, tiles, 512, first)
        assert plan[0][0] == plan[0][1] == first and plan[-1][2] == n
        for (i0, a0, b0), (i1, a1, b1) in zip(plan, plan[1:]):
            assert b0 == a1 and (a1 - first) % hop == 0 and (i1 - first) % hop == 0
            assert i1 == max(first, a1 - 512 * hop)
    with pytest.raises(ValueError):
        time_tiles(100, 0, 1)
