"""Read sharding across GPUs: the only multi-GPU structure of the query path.

Reads are independent, the index is not shardable (an LF step can land anywhere in the table), so every GPU
holds a replica of the table and takes one contiguous slice of the reads, balanced by bases (the reference
decomposes the same way over OpenMP threads: src/movi.cpp:274-301).  No collective on the query path; results
are concatenated in read order by whoever needs them.
"""
import numpy as np


def shard_bounds(offsets, world):
    """Contiguous read ranges [i0, i1) per rank with (nearly) equal numbers of bases."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    total = int(offsets[-1]) if n > 0 else 0
    cuts = [0]
    for k in range(1, world):
        target = total * k // world
        i = int(np.searchsorted(offsets, target, side="left"))
        cuts.append(min(max(i, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


def local_shard(bases, offsets, rank, world):
    """(bases_slice, offsets_rebased, (i0, i1)) of this rank."""
    i0, i1 = shard_bounds(offsets, world)[rank]
    o0, o1 = int(offsets[i0]), int(offsets[i1])
    return bases[o0:o1], (np.asarray(offsets[i0:i1 + 1], dtype=np.uint64) - np.uint64(o0)), (i0, i1)


def gather_in_read_order(local, dist=None, dtype=None):
    """all_gather variable-length per-rank result arrays and concatenate them in rank (= read) order.
    Works with any initialised torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, np.asarray(local))
    return np.concatenate(parts) if dtype is None else np.concatenate(parts).astype(dtype)
