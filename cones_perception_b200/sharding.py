"""Frame sharding across GPUs and the gather of per-frame cone lists.

Frames are independent on the GPU path, so a batch is cut into contiguous blocks of frames,
one block per rank, with NO collective on the data path.  The only exchange is the result
path: every rank's packed cone list (16 B per cone) plus its per-frame offsets are gathered
to rank 0.  Two implementations of the same packed format:
  * peer memory (production, `setup_peer_gather`): libconesgpu's publish kernel stores the list
    straight into rank 0's buffer over NVLink (CUDA-IPC mapping) at the end of every run;
  * one fixed-capacity all_gather (`gather_cone_lists`): NCCL on GPUs, gloo in the CPU tests.
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


def offset_words(frames_per_rank: int) -> int:
    """Words reserved for the cluster offsets in the packed result (F+1 rounded up to 4, the layout
    of libconesgpu's result block, see cp_device_results)."""
    return (frames_per_rank + 1 + 3) // 4 * 4


def pack_words(frames_per_rank: int, cap_cones: int) -> int:
    """int32 words of one rank's packed result: cluster offsets then cap_cones records."""
    return offset_words(frames_per_rank) + cap_cones * CONE_WORDS


def gather_cone_lists(packed, group=None, dst: int = 0):
    """Gather every rank's packed cone list to rank `dst` with ONE fixed-capacity all_gather.

    packed: int32 tensor [pack_words(F, cap)] on the collective's device: the rank's cluster offsets
            (cp_batch_results' cluster_offsets / cp_device_results) followed by the bit patterns of
            its cp_cluster records (the first offsets[-1] are valid).
    Returns on dst a tensor [world, F+1 + cap*4]; elsewhere None.  The message is KB-scale:
    latency-bound, not bandwidth-bound, so one collective beats a count exchange + gatherv.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    packed = packed.contiguous()
    out = torch.empty((world * packed.shape[0],), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    if rank != dst:
        return None
    return out.view(world, -1)


def setup_peer_gather(gpu, rank: int, world: int, frames_per_rank: int, cap_cones: int, group=None) -> int:
    """Wire the peer-memory result path of libconesgpu (cp_gather_*): rank 0 allocates the gather
    buffer and broadcasts its CUDA-IPC handle; the other ranks map it over NVLink.  After this every
    cp_batch_run publishes the rank's packed cone list into rank 0's memory from inside the library's
    own kernel — no collective on the step path.  Returns slot_words."""
    import torch
    import torch.distributed as dist

    slot_words = pack_words(frames_per_rank, cap_cones)
    dev = torch.device("cuda", torch.cuda.current_device())
    buf = torch.zeros(64, dtype=torch.uint8, device=dev)
    if rank == 0:
        handle = gpu.gather_create(world, slot_words)
        buf.copy_(torch.frombuffer(bytearray(handle), dtype=torch.uint8))
    dist.broadcast(buf, src=0, group=group)
    if rank != 0:
        gpu.gather_open(bytes(buf.cpu().numpy().tobytes()), rank, world, slot_words)
    dist.barrier(group=group)
    return slot_words


def unpack_gathered(gathered: np.ndarray, frames_per_rank: int):
    """Host-side view of a gather: list over global frames (rank-major) of structured cone arrays."""
    from .api import CLUSTER_DTYPE

    out = []
    g = np.ascontiguousarray(gathered).view(np.int32)
    for r in range(g.shape[0]):
        off = g[r, :frames_per_rank + 1].astype(np.int64)
        recs = g[r, offset_words(frames_per_rank):].view(np.uint32).reshape(-1, CONE_WORDS)
        for f in range(frames_per_rank):
            out.append(recs[off[f]:off[f + 1]].copy().view(CLUSTER_DTYPE).reshape(-1))
    return out
