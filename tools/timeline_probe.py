"""How do consecutive batches overlap when two handles alternate?  Device timestamps of the last two runs.
    python tools/timeline_probe.py [frames] [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cones_perception_b200 import api, scans  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 21
cfg = scans.config(3)
N = cfg.points_per_frame
dev = torch.from_numpy(scans.generate(cfg, F, 0)).cuda()
fp = np.full(F, N, np.uint32)
hs = []
for _ in range(2):
    h = api.ConesGpu(max_points=F * N, max_frames=F, max_survivors=F * N // 8, max_voxels=F * N // 16)
    h.set_device_input(dev.data_ptr(), fp, keep=dev)
    h.set_stage_timing(True)
    for _ in range(3):
        h.run(cfg.detect, cfg.ground)
    h.sync()
    hs.append(h)
names = ["run start", "pass1 start", "pass1 end", "pass2 start", "pass2 end", "run end"]
for lanes in (1, 2):
    torch.cuda.synchronize()
    for i in range(steps):
        hs[i % lanes].run(cfg.detect, cfg.ground)
    for h in hs:
        h.sync()
    last = hs[(steps - 1) % lanes]
    prev = hs[(steps - 2) % lanes]
    print(f"--- {lanes} lane(s): timestamps in ms after the start of the second-to-last run")
    if lanes == 2:
        print("second-to-last run:", "  ".join(f"{n} {t:.3f}" for n, t in zip(names, prev.timeline(prev))))
        print("last run          :", "  ".join(f"{n} {t:.3f}" for n, t in zip(names, last.timeline(prev))))
    else:
        print("last run          :", "  ".join(f"{n} {t:.3f}" for n, t in zip(names, last.timeline())))
