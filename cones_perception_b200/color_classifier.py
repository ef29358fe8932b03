"""Host-side mirror of the reference's colour service (`scripts/color_classifier_server.py`), with the work on
the device: `ColorClassifier.classify` is `handle_classify_color` (:78-126) for the cones of one frame.

The reference ships each cone's crop through a ROS service to a Python process that rasterises it (`to_image`,
:130-156) and runs `models/dam_net/dam_net.tflite` in the TFLite interpreter (:108-114).  Here the crop, the range
image and the network run on the cloud the detection call already left in HBM (`cp_cone_colors`); one byte per
cone comes back.  Response semantics are the service's:
  * a cone whose crop is empty is skipped — the response is SHORTER than the request (:83-84), so later colours
    shift forward (the node copies the response over the first entries of its list, src/cone_detection.cpp:357);
  * if `to_image` would raise for any cone (IndexError / interp1d ValueError), the service call fails as a whole
    and the node keeps every cone "unknown" (:356-361): `classify` then returns an empty list.
"""
from __future__ import annotations

import numpy as np

from . import api

COLORS = (None, "yellow", "blue", "orange")     # scripts/color_classifier_server.py:74


class ColorClassifier:
    def __init__(self, gpu: "api.ConesGpu", model_path, threshold: float = 0.8):
        """model_path: the launch parameter ~model_path (a .tflite file, :44-66) or its bytes."""
        self.gpu = gpu
        gpu.color_net_load_tflite(model_path, threshold)

    def classify(self, centers, cone_width: float = 0.228, msg=None, frame: int = 0) -> list[int]:
        """ClassifyColorSrvResponse.colors for the cones at `centers` (after the radial extension, :309):
        0 unknown, 1 yellow, 2 blue, 3 orange."""
        if len(centers) == 0:
            return []
        colors, _, flags = self.gpu.cone_colors(centers, cone_width, msg, frame)
        if np.any(flags & (api.CONE_BAD_INDEX | api.CONE_BAD_INTENSITY)):
            return []                                  # the service callback raises: no response at all
        return [int(c) for c, f in zip(colors, flags) if not (f & api.CONE_EMPTY)]
