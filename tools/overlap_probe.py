"""Experiment: do consecutive batches overlap when two handles (two streams) alternate?
    python tools/overlap_probe.py [frames] [runs]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cones_perception_b200 import api, scans  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 512
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cfg = scans.config(3)
N = cfg.points_per_frame
host = scans.generate(cfg, F, 0)
dev = torch.from_numpy(host).cuda()
fp = np.full(F, N, np.uint32)


def make():
    h = api.ConesGpu(max_points=F * N, max_frames=F, max_survivors=F * N // 8, max_voxels=F * N // 16)
    h.set_device_input(dev.data_ptr(), fp, keep=dev)
    for _ in range(3):
        h.run(cfg.detect, cfg.ground)
    h.sync()
    return h


def timed(handles):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(runs):
        handles[i % len(handles)].run(cfg.detect, cfg.ground)
    for h in handles:
        h.sync()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / runs * 1e3


a, b = make(), make()
print(f"streaming CTAs/SM cap {os.environ.get('CONESGPU_STREAM_CTAS', '8')}: one handle {timed([a]):.4f} ms/step, "
      f"two alternating handles {timed([a, b]):.4f} ms/step")
