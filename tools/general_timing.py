"""Configs 4 and 5 through the general back half: p50 ms per frame (single and batch of 8), launches, pairs tested.
    python tools/general_timing.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cones_perception_b200 import api, scans  # noqa: E402

for idx in (4, 5):
    cfg = scans.config(idx)
    for F in (1, 8):
        fr = scans.generate_config5(F, 0) if idx == 5 else scans.generate(cfg, F, 0)
        N = fr.shape[1]
        dev = torch.from_numpy(np.ascontiguousarray(fr)).cuda()
        with api.ConesGpu(max_points=F * N, max_frames=F) as g:
            g.set_device_input(dev.data_ptr(), np.full(F, N, np.uint32), keep=dev)
            for _ in range(4):
                g.run(cfg.detect, cfg.ground)
                g.sync()
            ctr, off, cl = g.results()
            lat = []
            for _ in range(30):
                torch.cuda.synchronize()
                t = time.perf_counter()
                g.run(cfg.detect, cfg.ground)
                g.sync()
                lat.append(1e3 * (time.perf_counter() - t))
            vis, tst = g.last_pairs()
            print(f"cfg{idx} F={F}: {np.percentile(lat, 50) / F:.3f} ms/frame  launches={g.last_launch_count()} "
                  f"C={int(ctr['n_cropped'][0])} V={int(ctr['n_voxels'][0])} K={int(off[1])} pairs visited={vis} tested={tst}")
