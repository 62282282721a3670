"""Multi-GPU sharding of the front-end: the path partitions by SEQUENCE.

Detection, description and stereo matching are per-frame independent and temporal KLT depends only on frame
t-1 of the same sequence, so the unit of work is a whole stereo sequence -- or a contiguous chunk of one, given a
one-frame overlap so that the chunk's first frame has its temporal predecessor.  One process per GPU owns its
shards; there is NO data-path collective.  ``torch.distributed`` is used only to agree on totals and timing
(`all_reduce` of a few scalars), which works on gloo (CPU tests) and NCCL alike.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Chunk:
    """frames [begin, end) of sequence `sequence`; `overlap` leading frames are re-processed only to seed the
    temporal tracks (their results belong to the previous chunk's owner and are dropped)."""
    sequence: int
    begin: int
    end: int
    overlap: int = 0

    @property
    def frames(self) -> int:
        """frames whose results this chunk owns"""
        return self.end - self.begin - self.overlap


def shard_sequences(lengths: list[int], world_size: int, rank: int) -> list[Chunk]:
    """Whole sequences to ranks, longest first onto the least loaded rank (deterministic, no communication:
    every rank computes the same assignment)."""
    assert 0 <= rank < world_size
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    load = [0] * world_size
    mine: list[Chunk] = []
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        load[r] += lengths[i]
        if r == rank:
            mine.append(Chunk(i, 0, lengths[i]))
    return sorted(mine, key=lambda c: c.sequence)


def shard_frames(total_frames: int, world_size: int, rank: int, sequence: int = 0) -> Chunk:
    """One long sequence as `world_size` contiguous chunks (BASELINE config 4: 4 096 frames over 8 GPUs).
    Every chunk but the first starts one frame early: that frame is detected again so that the temporal KLT
    jobs of the chunk's first owned frame have their previous-frame keypoints (results are bit-identical to the
    unsharded run because detection does not depend on history in this path)."""
    assert 0 <= rank < world_size
    base, rem = divmod(total_frames, world_size)
    begin = rank * base + min(rank, rem)
    end = begin + base + (1 if rank < rem else 0)
    overlap = 1 if (rank > 0 and end > begin) else 0
    return Chunk(sequence, begin - overlap, end, overlap)


def batches(chunk: Chunk, batch: int):
    """(first_frame, count) windows covering the chunk in order; the last window may be short"""
    f = chunk.begin
    while f < chunk.end:
        n = min(batch, chunk.end - f)
        yield f, n
        f += n


def global_sum(values: list[float]) -> list[float]:
    """sum of per-rank scalars over the default process group (identity when not initialised)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return list(values)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def global_max(value: float) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return value
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
