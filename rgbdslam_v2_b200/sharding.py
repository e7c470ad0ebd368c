"""Host-side sharding of frame pairs over ranks -- one process per GPU.

Frame pairs are independent, so rank r takes a contiguous range of the global pair list and runs
``match_node_pairs(..., first_pair_index=range.start)``: the counter-based random streams are keyed by the GLOBAL
pair index, so the union of the shards is bit-identical to a single-GPU run.  The only exchange is the all-gather
of the fixed-size edge records (NCCL in the C library: rgbdslam_b200_allgather_edges; torch.distributed / gloo here
for the CPU test of the host logic)."""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced ranges: the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def padded_shard_size(n_items: int, world: int) -> int:
    return -(-n_items // world)


def pad_edges(local: np.ndarray, size: int) -> np.ndarray:
    """Pad a rank's edge records to the common size with invalid edges (id1 = id2 = -1, node.cpp:1420)."""
    out = np.zeros(size, local.dtype)
    out["id1"] = -1
    out["id2"] = -1
    out[:len(local)] = local
    return out


def merge_gathered(all_edges: np.ndarray, n_items: int, world: int) -> np.ndarray:
    """Undo the padding of a rank-major all-gather: returns the n_items records in global pair order."""
    size = padded_shard_size(n_items, world)
    parts = [all_edges[r * size: r * size + len(shard_range(n_items, r, world))] for r in range(world)]
    return np.concatenate(parts)


def allgather_edges_torch(local: np.ndarray, n_items: int, world: int) -> np.ndarray:
    """torch.distributed implementation of the exchange (gloo on CPU, any backend): used by the CPU tests."""
    import torch
    import torch.distributed as dist
    size = padded_shard_size(n_items, world)
    buf = torch.from_numpy(pad_edges(local, size).view(np.uint8).copy())
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    allb = np.concatenate([o.numpy() for o in out]).view(local.dtype)
    return merge_gathered(allb, n_items, world)


def frame_shard(total_frames: int, rank: int, world: int) -> range:
    """Frames of a sequence owned by `rank` in rgbdslam_b200_nodes_create_sharded: blocks of ceil(total / world)."""
    per = -(-total_frames // world)
    return range(min(rank * per, total_frames), min((rank + 1) * per, total_frames))
