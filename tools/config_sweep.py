"""All five BASELINE.json configurations, one frame each (and a batch of 8 for 4 and 5): device-resident
time per frame through cp_batch_run, the back-half variant that ended up running, and the CPU restatement
on one core beside it.    python tools/config_sweep.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cones_perception_b200 import api, scans  # noqa: E402
from oracle import oracle as O  # noqa: E402


def frames_of(idx, n, seed=0):
    return scans.generate_config5(n, seed) if idx == 5 else scans.generate(scans.config(idx), n, seed)


for idx in (1, 2, 3, 4, 5):
    cfg = scans.config(idx)
    for F in ((1, 8) if idx in (4, 5) else (1,)):
        fr = frames_of(idx, F)
        N = fr.shape[1]
        dev = torch.from_numpy(np.ascontiguousarray(fr)).cuda()
        with api.ConesGpu(max_points=F * N, max_frames=F) as g:
            g.set_device_input(dev.data_ptr(), np.full(F, N, np.uint32), keep=dev)
            for _ in range(3):
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
            ms = float(np.percentile(lat, 50))
            launches = g.last_launch_count()
        t = time.perf_counter()
        ecl, ectr, _ = O.detect(O.view_of_xyzi(fr[0]), cfg.detect, cfg.ground, O.PCL_FAITHFUL)
        cpu_ms = 1e3 * (time.perf_counter() - t)
        c0 = ctr[0]
        print(f"cfg{idx} F={F}: N={N} C={int(c0['n_cropped'])} V={int(c0['n_voxels'])} comps={int(c0['n_components'])} "
              f"K={int(off[1])} | GPU {ms / F:.3f} ms/frame ({F * N / ms / 1e6:.2f} G pts/s), {launches} launches | "
              f"CPU one core {cpu_ms:.1f} ms/frame | x{cpu_ms / (ms / F):.0f}")
