"""Frames of config 4 or 5 through the general back half, for profiling:
    python tools/general_case.py [reps] [config] [frames]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cones_perception_b200 import api, scans  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 5
F = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = scans.config(idx)
fr = scans.generate_config5(F, 0) if idx == 5 else scans.generate(cfg, F, 0)
N = fr.shape[1]
dev = torch.from_numpy(np.ascontiguousarray(fr)).cuda()
with api.ConesGpu(max_points=F * N, max_frames=F, env={"CONESGPU_GRAPH": "0"}) as g:
    g.set_device_input(dev.data_ptr(), np.full(F, N, np.uint32), keep=dev)
    for _ in range(reps):
        g.run(cfg.detect, cfg.ground)
        g.sync()
    print("launches", g.last_launch_count())
