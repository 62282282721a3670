"""Host-side sharding logic, including a world-size-2 gloo run on CPU: the N>1 path has no data-path
collective, so what needs covering is (a) every frame is owned exactly once, (b) chunk overlap seeds the temporal
tracks, (c) the scalar reductions bench.py relies on work across ranks."""
import os
import socket

import numpy as np
import pytest

from zenslam_b200 import sharding


def test_shard_sequences_partition():
    lengths = [100, 7, 64, 64, 3, 250, 31]
    for world in (1, 2, 3, 8):
        seen = []
        loads = []
        for r in range(world):
            mine = sharding.shard_sequences(lengths, world, r)
            seen += [c.sequence for c in mine]
            loads.append(sum(c.frames for c in mine))
            assert all(c.begin == 0 and c.end == lengths[c.sequence] and c.overlap == 0 for c in mine)
        assert sorted(seen) == list(range(len(lengths)))
        assert sum(loads) == sum(lengths)
        assert max(loads) <= max(max(lengths), -(-sum(lengths) // world) + max(lengths))


@pytest.mark.parametrize("total,world", [(4096, 8), (10, 3), (5, 8), (0, 2), (1, 1)])
def test_shard_frames_cover_once_with_overlap(total, world):
    owned = np.zeros(total, np.int32)
    for r in range(world):
        c = sharding.shard_frames(total, world, r)
        assert c.frames >= 0
        first_owned = c.begin + c.overlap
        owned[first_owned:c.end] += 1
        if r > 0 and c.frames > 0:
            assert c.overlap == 1 and c.begin == first_owned - 1      # previous frame is re-processed, not owned
        got = sum(n for _, n in sharding.batches(c, 16))
        assert got == c.end - c.begin
    assert np.all(owned == 1)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = sharding.shard_frames(101, world, rank)
        frames, = sharding.global_sum([float(c.frames)])
        slowest = sharding.global_max(10.0 + rank)
        q.put((rank, c.begin, c.end, c.overlap, frames, slowest))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, b0, e0, o0, f0, s0), (r1, b1, e1, o1, f1, s1) = out
    assert (b0, e0, o0) == (0, 51, 0) and (b1, e1, o1) == (50, 101, 1)
    assert f0 == f1 == 101.0                 # every frame owned exactly once across the two ranks
    assert s0 == s1 == 11.0                  # max-over-ranks, as bench.py reports time
