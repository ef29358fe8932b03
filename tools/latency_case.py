"""Single-frame latency case (config 2) for profiling: `python tools/latency_case.py [reps]`."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cones_perception_b200 import api, scans  # noqa: E402
from cones_perception_b200.pointcloud2 import PointCloud2  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
cfg = scans.config(2)
f = scans.generate(cfg, 1, base_seed=0)[0]
pin = torch.empty(f.shape, dtype=torch.float32, pin_memory=True)
pin.numpy()[:] = f
msg = PointCloud2.from_xyzi(pin.numpy())
with api.ConesGpu(max_points=len(f), max_frames=1) as gpu:
    for _ in range(5):
        gpu.detect(msg, cfg.detect, cfg.ground)
    lat = []
    for _ in range(reps):
        t = time.perf_counter()
        cl, ctr = gpu.detect(msg, cfg.detect, cfg.ground)
        lat.append(1e6 * (time.perf_counter() - t))
    print(f"pinned   p50 {np.percentile(lat, 50):.1f} us  p99 {np.percentile(lat, 99):.1f} us  K={len(cl)} launches={gpu.last_launch_count()}")
    pmsg = PointCloud2.from_xyzi(f.copy())      # pageable host memory, like a ROS message's std::vector
    lat = []
    for _ in range(reps):
        t = time.perf_counter()
        cl, ctr = gpu.detect(pmsg, cfg.detect, cfg.ground)
        lat.append(1e6 * (time.perf_counter() - t))
    print(f"pageable p50 {np.percentile(lat, 50):.1f} us  p99 {np.percentile(lat, 99):.1f} us")

# the same call with every ctypes argument prepared once (what a C++ node pays): how much of the figures above
# is the Python binding's per-call work?
import ctypes as C  # noqa: E402
from cones_perception_b200.params import to_c_detect, to_c_ground  # noqa: E402
from cones_perception_b200.pointcloud2 import make_view  # noqa: E402
with api.ConesGpu(max_points=len(f), max_frames=1) as gpu:
    view = make_view(msg, True)
    cd, cg = to_c_detect(cfg.detect), to_c_ground(cfg.ground)
    out = np.zeros(4096, dtype=api.CLUSTER_DTYPE)
    ctr = np.zeros(1, dtype=api.COUNTER_DTYPE)
    k = C.c_uint32()
    args = (gpu._h, C.byref(view), C.byref(cd), C.byref(cg), out.ctypes.data, 4096, C.byref(k), ctr.ctypes.data)
    fn = gpu.lib.cp_detect
    for _ in range(5):
        fn(*args)
    lat = []
    for _ in range(reps):
        t = time.perf_counter()
        fn(*args)
        lat.append(1e6 * (time.perf_counter() - t))
    print(f"pinned, C ABI with prepared arguments: p50 {np.percentile(lat, 50):.1f} us  p99 {np.percentile(lat, 99):.1f} us  K={k.value}")
