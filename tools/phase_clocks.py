import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from cones_perception_b200 import api, scans
from cones_perception_b200.pointcloud2 import PointCloud2
cfg = scans.config(3)
F = int(sys.argv[1])
fr = scans.generate(cfg, F, 0)
with api.ConesGpu(max_points=F*cfg.points_per_frame, max_frames=F) as g:
    msgs=[PointCloud2.from_xyzi(f) for f in fr]
    g.set_host_input(msgs); g.run(cfg.detect, cfg.ground); g.sync()
    print("---- second run"); sys.stdout.flush()
    g.run(cfg.detect, cfg.ground); g.sync()
