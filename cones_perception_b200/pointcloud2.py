"""A ROS-free stand-in for sensor_msgs/PointCloud2 (the message both reference callbacks take,
src/cone_detection.cpp:130, src/ground_removal.cpp:50) plus the field resolution that
pcl::fromROSMsg performs (exact name / FLOAT32 / count 1 match)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

FLOAT32 = 7  # sensor_msgs/PointField.FLOAT32


@dataclass
class PointField:
    name: str
    offset: int
    datatype: int = FLOAT32
    count: int = 1


@dataclass
class PointCloud2:
    data: np.ndarray                      # uint8, C-contiguous, height*row_step bytes
    width: int
    height: int = 1
    point_step: int = 16
    row_step: int = 0
    fields: list = field(default_factory=list)
    is_bigendian: bool = False
    is_dense: bool = True
    frame_id: str = "cloud"
    stamp: tuple = (0, 0)                 # (sec, nsec)

    def __post_init__(self):
        if self.row_step == 0:
            self.row_step = self.width * self.point_step

    @staticmethod
    def from_xyzi(xyzi: np.ndarray, with_intensity: bool = True) -> "PointCloud2":
        """Compact cloud: float32 x,y,z,intensity at offsets 0,4,8,12, point_step 16."""
        a = np.ascontiguousarray(xyzi, dtype=np.float32).reshape(-1, 4)
        f = [PointField("x", 0), PointField("y", 4), PointField("z", 8)]
        if with_intensity:
            f.append(PointField("intensity", 12))
        return PointCloud2(data=a.view(np.uint8).reshape(-1), width=a.shape[0], height=1, point_step=16, fields=f)

    def offset_of(self, name: str) -> int:
        for f in self.fields:
            if f.name == name and f.datatype == FLOAT32 and f.count == 1:
                return f.offset
        return -1

    @property
    def n_points(self) -> int:
        return self.width * self.height


class CCloudView(C.Structure):
    """cp_cloud_view / orc_view (identical layout)."""
    _fields_ = [
        ("data", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32),
        ("point_step", C.c_uint32), ("row_step", C.c_uint32),
        ("off_x", C.c_int32), ("off_y", C.c_int32), ("off_z", C.c_int32), ("off_intensity", C.c_int32),
        ("is_bigendian", C.c_uint8), ("is_dense", C.c_uint8),
    ]


def make_view(msg: PointCloud2, fake_missing_intensity: bool) -> CCloudView:
    """Resolve fields like pcl::fromROSMsg.  ``fake_missing_intensity`` reproduces
    src/cone_detection.cpp:142-151: a missing ``intensity`` field is faked at offset 0."""
    oi = msg.offset_of("intensity")
    if oi < 0 and fake_missing_intensity:
        oi = 0
    data = msg.data
    if not (isinstance(data, np.ndarray) and data.dtype == np.uint8 and data.flags["C_CONTIGUOUS"]):
        raise TypeError("PointCloud2.data must be a C-contiguous uint8 numpy array")
    return CCloudView(data.ctypes.data, msg.width, msg.height, msg.point_step, msg.row_step,
                      msg.offset_of("x"), msg.offset_of("y"), msg.offset_of("z"), oi,
                      1 if msg.is_bigendian else 0, 1 if msg.is_dense else 0)
