"""Frame sharding across GPUs and the gather of per-frame cone lists.

Frames are independent on the GPU path, so a batch is cut into contiguous blocks of frames,
one block per rank, with NO collective on the data path.  The only exchange is the result
path: every rank's packed cone list (16 B per cone) plus its per-frame counts are gathered
to rank 0 — NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

CONE_WORDS = 4  # cp_cluster = 4 x 32-bit words


def shard_frames(n_frames: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of frames for `rank` (blocks differ by at most one frame and
    keep frame order, so the sequential host-side temporal gate can run per rank)."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_cone_lists(counts, cones, cap_cones: int, group=None, dst: int = 0):
    """Gather per-frame cone counts and packed cone records to rank `dst`.

    counts: int32 tensor [frames_per_rank] (cones per local frame), on the collective's device
    cones:  int32 tensor [cap_cones, 4] (bit patterns of cp_cluster records, first sum(counts) valid)
    Returns on dst: (counts_all [world, frames_per_rank], cones_all [world, cap_cones, 4]);
    elsewhere (None, None).  Fixed-capacity all_gather keeps it to two collectives of KB-size
    messages: latency-bound, not bandwidth-bound.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = counts.contiguous()
    cones = cones.contiguous()
    assert cones.shape == (cap_cones, CONE_WORDS)
    # outputs are the dim-0 concatenation of the inputs (the form both NCCL and gloo accept)
    counts_all = torch.empty((world * counts.shape[0],), dtype=counts.dtype, device=counts.device)
    cones_all = torch.empty((world * cap_cones, CONE_WORDS), dtype=cones.dtype, device=cones.device)
    dist.all_gather_into_tensor(counts_all, counts, group=group)
    dist.all_gather_into_tensor(cones_all, cones, group=group)
    if rank != dst:
        return None, None
    return counts_all.view(world, -1), cones_all.view(world, cap_cones, CONE_WORDS)


def unpack_gathered(counts_all: np.ndarray, cones_all: np.ndarray):
    """Host-side view of a gather: list over global frames of structured cone arrays."""
    from .api import CLUSTER_DTYPE

    out = []
    world = counts_all.shape[0]
    for r in range(world):
        recs = np.ascontiguousarray(cones_all[r]).view(np.uint32).reshape(-1, CONE_WORDS)
        off = 0
        for c in counts_all[r]:
            c = int(c)
            out.append(recs[off:off + c].copy().view(CLUSTER_DTYPE).reshape(-1))
            off += c
    return out
