"""Small end-to-end case for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py
Runs both back-half variants, a ragged batch, the node-equivalent ground removal and checks
every result against the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cones_perception_b200 import api, scans  # noqa: E402
from cones_perception_b200.params import GroundParams  # noqa: E402
from cones_perception_b200.pointcloud2 import PointCloud2  # noqa: E402
from oracle import oracle as O  # noqa: E402

cfg = scans.config(1)
frames = list(scans.generate(cfg, 3, base_seed=1))
frames[1] = frames[1][:5000]
for mode in (0, 3):
    with api.ConesGpu(max_points=1 << 17, max_frames=4, back_mode=mode) as gpu:
        ctr, off, cl = gpu.detect_batch([PointCloud2.from_xyzi(f) for f in frames], cfg.detect, GroundParams())
        for i, f in enumerate(frames):
            exp, _, _ = O.detect(O.view_of_xyzi(np.ascontiguousarray(f)), cfg.detect, GroundParams(), O.CANONICAL)
            got = cl[off[i]:off[i + 1]]
            assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (mode, i)
        out, kept, low = gpu.ground_remove(PointCloud2.from_xyzi(frames[0]), GroundParams())
        _, ekept, elow, _ = O.ground_node(O.view_of_xyzi(frames[0]), GroundParams())
        assert kept == ekept and np.array_equal(low, elow)
print("sanitize case ok")
