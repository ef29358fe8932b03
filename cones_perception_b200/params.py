"""Parameter surface of the reference nodes (config/*.yaml), names kept verbatim —
including the reference's spelling (``distance_treshold_*``, ``cones_matching_dist_theshold``).

Defaults are the member initialisers of ``ConeDetector`` (src/cone_detection.cpp:22-43) and
``GroundRemover`` (src/ground_removal.cpp:18-19); the presets are the three shipped files
config/cones_detection_params_{our,fsai,simulation}.yaml and config/ground_removal_params.yaml.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import asdict, dataclass


@dataclass
class DetectParams:
    distance_treshold_max: float = 7.0
    distance_treshold_min: float = 0.7
    level_threshold: float = -0.5
    angle_threshold: float = 90.0
    voxel_filter_leaf_size_x: float = 0.04
    voxel_filter_leaf_size_y: float = 0.04
    voxel_filter_leaf_size_z: float = 0.04
    min_cluster_size: int = 3
    max_cluster_size: int = 50
    cones_matching_dist_theshold: float = 0.5
    cone_position_extension_length: float = 0.05
    use_points_buffer: bool = False
    classify_colors: bool = False
    # const members, src/cone_detection.cpp:22-23
    CONE_WIDTH: float = 0.228
    CONE_HEIGHT: float = 0.325


@dataclass
class GroundParams:
    num_of_sectors: int = 16
    default_lowest_point: float = -0.1


PRESETS = {
    # config/cones_detection_params_our.yaml:1-11
    "our": DetectParams(7.0, 0.7, -0.5, 90.0, 0.04, 0.04, 0.04, 3, 50, 0.5, 0.05),
    # config/cones_detection_params_fsai.yaml:1-11
    "fsai": DetectParams(6.0, 1.0, -0.09, 160.0, 0.04, 0.04, 0.04, 3, 500, 0.5, 0.05),
    # config/cones_detection_params_simulation.yaml:1-11
    "simulation": DetectParams(10.0, 1.0, -5.0, 160.0, 0.04, 0.04, 0.04, 2, 500, 0.5, 0.05),
}


def load_yaml_params(path: str, base: DetectParams | GroundParams | None = None):
    """Load one of the reference's yaml files onto the node defaults (ros::param::get semantics:
    keys that are absent keep the member default)."""
    import yaml

    with open(path) as f:
        data = yaml.safe_load(f) or {}
    if base is None:
        base = GroundParams() if ("num_of_sectors" in data or "default_lowest_point" in data) else DetectParams()
    out = type(base)(**asdict(base))
    for k, v in data.items():
        if hasattr(out, k):
            setattr(out, k, type(getattr(out, k))(v))
    return out


class CDetectParams(C.Structure):
    """cp_detect_params / orc_detect_params (identical layout)."""
    _fields_ = [
        ("distance_treshold_max", C.c_double), ("distance_treshold_min", C.c_double),
        ("level_threshold", C.c_double), ("angle_threshold", C.c_double),
        ("voxel_filter_leaf_size_x", C.c_double), ("voxel_filter_leaf_size_y", C.c_double),
        ("voxel_filter_leaf_size_z", C.c_double),
        ("min_cluster_size", C.c_int32), ("max_cluster_size", C.c_int32),
        ("cone_width", C.c_float), ("cone_height", C.c_float),
    ]


class CGroundParams(C.Structure):
    _fields_ = [("num_of_sectors", C.c_int32), ("default_lowest_point", C.c_float)]


def to_c_detect(p: DetectParams) -> CDetectParams:
    return CDetectParams(p.distance_treshold_max, p.distance_treshold_min, p.level_threshold, p.angle_threshold,
                         p.voxel_filter_leaf_size_x, p.voxel_filter_leaf_size_y, p.voxel_filter_leaf_size_z,
                         int(p.min_cluster_size), int(p.max_cluster_size), p.CONE_WIDTH, p.CONE_HEIGHT)


def to_c_ground(p: GroundParams) -> CGroundParams:
    return CGroundParams(int(p.num_of_sectors), float(p.default_lowest_point))
