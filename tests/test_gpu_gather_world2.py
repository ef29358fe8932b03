"""The peer-memory result path (cp_gather_*) with TWO processes on TWO GPUs: rank 1 publishes its cone lists into
rank 0's memory over a CUDA-IPC mapping; rank 0 must read exactly what rank 1's own cp_batch_results returned —
for several runs (both parities of the double buffer), with the data changing between runs so that a stale slot
cannot pass, and through a replayed CUDA graph.  Skipped when fewer than two GPUs are visible."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from cones_perception_b200 import api, scans
    from cones_perception_b200.pointcloud2 import PointCloud2
    from cones_perception_b200.sharding import pack_words, setup_peer_gather, unpack_gathered
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = scans.config(3)
    F, cap = 6, 6 * 64
    words = pack_words(F, cap)
    out = []
    with api.ConesGpu(max_points=F * cfg.points_per_frame, max_frames=F, device=rank) as h:
        setup_peer_gather(h, rank, world, F, cap)
        for run in range(5):
            # runs 0-2: fresh frames each time; runs 3-4 repeat run 2's input (direct -> graph capture -> replay)
            seed = 1000 * rank + 10 * min(run, 2)
            frames = list(scans.generate(cfg, F, base_seed=seed))
            h.set_host_input([PointCloud2.from_xyzi(f) for f in frames])
            h.run(cfg.detect, cfg.ground)
            _, off, cl = h.results()
            mine = [cl[off[f]:off[f + 1]].copy() for f in range(F)]
            everyone = [None] * world
            dist.all_gather_object(everyone, mine)
            if rank == 0:
                seq = h.gather_seq()
                assert seq == run + 1
                h.gather_wait(seq)
                got = unpack_gathered(h.gather_read(seq, world, words), F)
                for r in range(world):
                    for f in range(F):
                        a, b = got[r * F + f], everyone[r][f]
                        assert len(a) == len(b) and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (run, r, f)
                out.append(sum(len(x) for x in got))
            dist.barrier()
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


def test_peer_gather_two_processes_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with `gpurun --gpus 2`); the one-GPU box covers world = 1 only")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    totals = q.get(timeout=300)
    [p.join(timeout=120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert len(totals) == 5 and all(t > 0 for t in totals)
