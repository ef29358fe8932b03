"""CPU tests of the multi-GPU host logic: frame sharding and the cone-list gather, run with
world_size 2 over gloo (the GPU box uses the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest

from cones_perception_b200.sharding import shard_frames, unpack_gathered


@pytest.mark.parametrize("n,world", [(4096, 8), (4096, 1), (10, 3), (7, 8), (0, 4), (512, 2)])
def test_shards_are_contiguous_and_cover(n, world):
    blocks = [shard_frames(n, world, r) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in blocks]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from cones_perception_b200.api import CLUSTER_DTYPE
    from cones_perception_b200.sharding import gather_cone_lists, offset_words, pack_words
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames_per_rank, cap = 3, 16
    counts = np.array([(rank + 1), 0, 2 * (rank + 1)], np.int32)
    recs = np.zeros(cap, CLUSTER_DTYPE)
    k = int(counts.sum())
    recs["x"][:k] = 100 * rank + np.arange(k)
    recs["y"][:k] = -1.5
    recs["size"][:k] = 3 + np.arange(k)
    recs["min_index"][:k] = np.arange(k)
    packed = np.zeros(pack_words(frames_per_rank, cap), np.int32)
    packed[1:frames_per_rank + 1] = np.cumsum(counts)
    packed[offset_words(frames_per_rank):] = recs.view(np.int32)
    got = gather_cone_lists(torch.from_numpy(packed))
    if rank == 0:
        q.put(got.numpy())
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_cone_lists_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    gathered = q.get(timeout=120)
    [p.join(timeout=120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert gathered.shape == (2, 4 + 16 * 4)  # offsets padded to 4 words + 16 records
    frames = unpack_gathered(gathered, 3)
    assert [len(f) for f in frames] == [1, 0, 2, 2, 0, 4]
    assert frames[0]["x"].tolist() == [0.0] and frames[2]["x"].tolist() == [1.0, 2.0]
    assert frames[3]["x"].tolist() == [100.0, 101.0] and frames[5]["size"].tolist() == [5, 6, 7, 8]
    assert all((f["y"] == -1.5).all() for f in frames if len(f))
