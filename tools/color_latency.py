"""Colour-path inputs on one config-2 frame: `python tools/color_latency.py [reps]`.
Times cp_cone_crops / cp_cone_images on the cloud the detection call left on the device, next to the
CPU restatement of the reference's per-cone loop + numpy-equivalent raster (oracle, one core)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cones_perception_b200 import api, scans  # noqa: E402
from cones_perception_b200.pointcloud2 import PointCloud2  # noqa: E402
from oracle import oracle as O  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
cfg = scans.config(2)
f = scans.generate(cfg, 1, base_seed=0)[0]
msg = PointCloud2.from_xyzi(f)
with api.ConesGpu(max_points=len(f), max_frames=1) as gpu:
    cl, _ = gpu.detect(msg, cfg.detect, cfg.ground)
    centers = np.array([O.extend(float(c["x"]), float(c["y"]), 0.05) for c in cl], np.float32)
    for _ in range(5):
        gpu.cone_crops(centers)
        gpu.cone_images(centers)

    def p50(fn):
        lat = []
        for _ in range(reps):
            t = time.perf_counter()
            fn()
            lat.append(1e6 * (time.perf_counter() - t))
        return np.percentile(lat, 50), np.percentile(lat, 99)

    off, pts = gpu.cone_crops(centers)
    # the C ABI with caller-owned buffers (what a node does), without the Python wrapper's allocations
    import ctypes as C
    n = len(centers)
    o = np.zeros(n + 1, np.uint32)
    buf = np.empty((16384, 4), np.float32)
    img = np.zeros((n, 15, 12), np.uint8)
    cnt, flg = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    a = p50(lambda: gpu.lib.cp_cone_crops(gpu._h, None, 0, centers.ctypes.data, n, C.c_float(0.228), o.ctypes.data,
                                          buf.ctypes.data, 16384))
    b = p50(lambda: gpu.lib.cp_cone_images(gpu._h, None, 0, centers.ctypes.data, n, C.c_float(0.228),
                                           img.ctypes.data, cnt.ctypes.data, flg.ctypes.data))
    assert np.array_equal(o, off)
    print(f"{len(centers)} cones, {off[-1]} crop points of {len(f)}")
    print(f"cp_cone_crops  (device cloud -> host crops)   p50 {a[0]:.1f} us  p99 {a[1]:.1f} us")
    print(f"cp_cone_images (device cloud -> host images)  p50 {b[0]:.1f} us  p99 {b[1]:.1f} us")
cloud = O.from_msg(O.view_of_xyzi(f))
t = time.perf_counter()
n = 5
for _ in range(n):
    crops = [O.reconstruct_cone(cloud, float(c[0]), float(c[1]), 0.228) for c in centers]
t1 = (time.perf_counter() - t) / n
t = time.perf_counter()
for _ in range(n):
    for c in crops:
        O.to_image(np.stack([c["x"], c["y"], c["z"], c["intensity"]], 1))
t2 = (time.perf_counter() - t) / n
print(f"CPU restatement, one core: per-cone loop over the cloud {1e6 * t1:.0f} us, raster {1e6 * t2:.0f} us "
      "(C loops called from Python; the reference's raster is numpy, its transport a ROS service)")
