#!/usr/bin/env python
"""Raw host->device ceiling of the box, per rank and aggregate.

    python tools/h2d_probe.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/h2d_probe.py                  # N ranks copying at the same time

Each rank binds itself next to its GPU (cones_perception_b200/placement.py), allocates a 1 GiB pinned buffer and
runs a bare loop of cudaMemcpyAsync copies; rank 0 prints one JSON line with per-rank and aggregate GB/s, plus the
same loop with the ranks NOT bound (BIND=0) when --both is given.  bench.py runs the same probe inside its own
process group and reports e2e as a fraction of it.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=1 << 30)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--no-bind", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from cones_perception_b200 import placement

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {"world": world, "cpus_allowed": len(placement.allowed_cpus()),
           "numa_nodes_visible": {str(k): len(v) for k, v in placement.numa_nodes().items()},
           "gpu_numa_node": placement.gpu_numa_node(local)}
    # unbound first (the process has not been moved yet), then bound
    res = {}
    for mode in (["unbound"] if args.no_bind else ["unbound", "bound"]):
        bind = None
        if mode == "bound":          # ranks find their placement one after the other: the probes must not interfere
            for r in range(world):
                if r == rank:
                    bind = placement.bind_rank(local)
                barrier()
        barrier()
        pr = placement.h2d_probe(local, args.bytes, args.reps, barrier)
        rec = {"GBps": pr["GBps"], "pages_by_node": pr["pages_by_node"], "bind": bind}
        if world > 1:
            allrec = [None] * world
            dist.all_gather_object(allrec, rec)
        else:
            allrec = [rec]
        res[mode] = {"per_rank_GBps": [r["GBps"] for r in allrec], "aggregate_GBps": sum(r["GBps"] for r in allrec),
                     "ranks": allrec}
    out["h2d"] = res
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
