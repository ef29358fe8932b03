"""cones_perception_b200 — B200-native (sm_100a) point-cloud hot path of dmn-sjk/cones_perception.

The product is ``libconesgpu.so`` (hand-written CUDA behind the C ABI in ``include/conesgpu.h``).
This package holds its sources (``csrc/``), the build recipe, a ctypes binding that mirrors
the reference's two node classes (``GroundRemover`` / ``ConeDetector``, ROS-free), the
synthetic scan generator and the frame-sharding helpers for multi-GPU batches.
There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from .params import DetectParams, GroundParams, PRESETS, load_yaml_params  # noqa: F401
from .pointcloud2 import PointCloud2, PointField  # noqa: F401

__version__ = "0.1.0"
