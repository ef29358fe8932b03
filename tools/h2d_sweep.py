#!/usr/bin/env python
"""Where does the host->device ceiling of an N-GPU box come from?  Run under torch.distributed.run with one rank
per GPU.  Rank 0 prints one JSON document:
  solo        each rank copies alone (per-link rate)
  together_k  ranks 0..k-1 copy at the same time (k = 1, 2, 4, .. N): where the aggregate saturates
  variants    all N ranks together: default pinned / write-combined / two streams / after bind_rank()
Every copy is a plain pinned cudaMemcpyAsync of --bytes (default 1 GiB), --reps per measurement, CUDA events.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=1 << 30)
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from cones_perception_b200 import api, placement

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        subprocess.run(["bash", os.path.join(ROOT, "tools", "box_topology.sh"),
                        os.path.join(ROOT, "gpurun_out", f"box_topology_n{world}.txt")], check=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(v):
        if world == 1:
            return [v]
        out = [None] * world
        dist.all_gather_object(out, v)
        return out

    dev = torch.empty(args.bytes, dtype=torch.uint8, device="cuda")
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]

    def tensor_of(buf):
        return torch.from_numpy(buf.array)

    def timed(host, active: bool, nstreams: int = 1) -> float:
        """All ranks enter; only `active` ranks copy.  Returns this rank's GB/s (0 when idle)."""
        barrier()
        gb = 0.0
        if active:
            parts = [(host, dev)] if nstreams == 1 else \
                [(host[:args.bytes // 2], dev[:args.bytes // 2]), (host[args.bytes // 2:], dev[args.bytes // 2:])]
            e0 = torch.cuda.Event(enable_timing=True)
            ends = [torch.cuda.Event(enable_timing=True) for _ in parts]
            e0.record(streams[0])
            streams[1].wait_event(e0)
            for _ in range(args.reps):
                for (h, d), s in zip(parts, streams):
                    with torch.cuda.stream(s):
                        d.copy_(h, non_blocking=True)
            for e, s in zip(ends, streams):
                e.record(s)
            ms = 0.0
            for e in ends:
                e.synchronize()
                ms = max(ms, e0.elapsed_time(e))
            gb = args.reps * args.bytes / (ms * 1e-3) / 1e9
        barrier()
        return gb

    out = {"world": world, "bytes": args.bytes, "reps": args.reps, "cpus_allowed": len(placement.allowed_cpus()),
           "numa_nodes_visible": {str(k): len(v) for k, v in placement.numa_nodes().items()}}
    out["gpu_numa_node"] = gather(placement.gpu_numa_node(local))

    plain = api.PinnedBuffer(args.bytes, local, write_combined=False)
    plain.array[:] = 1
    hp = tensor_of(plain)
    timed(hp, True)                                     # warm-up
    out["pages_unbound"] = gather(placement.page_nodes(plain.ptr, args.bytes))
    out["solo"] = [max(gather(timed(hp, rank == r))) for r in range(world)]
    k = 1
    while k <= world:
        out[f"together_{k}"] = gather(timed(hp, rank < k))[:k]
        k *= 2
    variants = {"pinned_unbound": gather(timed(hp, True)), "pinned_unbound_2streams": gather(timed(hp, True, 2))}
    wc = api.PinnedBuffer(args.bytes, local, write_combined=True)
    wc.array[:] = 1
    hw = tensor_of(wc)
    timed(hw, True)
    variants["write_combined_unbound"] = gather(timed(hw, True))
    hw = None
    wc.close()
    hp = None
    plain.close()

    bind = None
    for r in range(world):                              # one rank at a time: the placement probes must not interfere
        if r == rank:
            bind = placement.bind_rank(local)
        barrier()
    out["bind"] = gather(bind)
    bound = api.PinnedBuffer(args.bytes, local, write_combined=False)
    bound.array[:] = 1
    hb = tensor_of(bound)
    timed(hb, True)
    out["pages_bound"] = gather(placement.page_nodes(bound.ptr, args.bytes))
    variants["pinned_bound"] = gather(timed(hb, True))
    variants["pinned_bound_2streams"] = gather(timed(hb, True, 2))
    out["variants"] = variants
    out["aggregate_GBps"] = {k: float(np.sum(v)) for k, v in variants.items()}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
