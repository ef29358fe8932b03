"""ctypes wrapper of the CPU oracle (oracle/cones_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never by the product package.  Pinned against the reference's own compiled node sources
(oracle/ref.py, oracle/_ref) for everything the reference wrote itself; PARITY UNPINNED for PCL's
VoxelGrid / EuclideanClusterExtraction, which are not available here (see cones_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libconesoracle.so")

CANONICAL, PCL_FAITHFUL = 0, 1
NSECT = 17

POINT_DTYPE = np.dtype([("x", np.float32), ("y", np.float32), ("z", np.float32), ("pad", np.float32),
                        ("intensity", np.float32), ("c1", np.float32), ("c2", np.float32), ("c3", np.float32)])
CLUSTER_DTYPE = np.dtype([("x", np.float32), ("y", np.float32), ("size", np.uint32), ("min_index", np.uint32)])


class View(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("point_step", C.c_uint32),
                ("row_step", C.c_uint32), ("off_x", C.c_int32), ("off_y", C.c_int32), ("off_z", C.c_int32),
                ("off_intensity", C.c_int32), ("is_bigendian", C.c_uint8), ("is_dense", C.c_uint8)]


class DetectParams(C.Structure):
    _fields_ = [("distance_treshold_max", C.c_double), ("distance_treshold_min", C.c_double),
                ("level_threshold", C.c_double), ("angle_threshold", C.c_double),
                ("voxel_filter_leaf_size_x", C.c_double), ("voxel_filter_leaf_size_y", C.c_double),
                ("voxel_filter_leaf_size_z", C.c_double), ("min_cluster_size", C.c_int32),
                ("max_cluster_size", C.c_int32), ("cone_width", C.c_float), ("cone_height", C.c_float)]


class GroundParams(C.Structure):
    _fields_ = [("num_of_sectors", C.c_int32), ("default_lowest_point", C.c_float)]


class Timing(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("from_msg", "copy_cloud", "ground", "crop", "voxel", "cluster", "centroid",
                                          "total")]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("n_points", "n_ground_kept", "n_cropped", "n_voxels", "n_components",
                                          "n_clusters", "key_bits", "passthrough")] + \
               [("min_b", C.c_int32 * 3), ("div_b", C.c_int32 * 3)]


_lib = None


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("cones_oracle.cpp", "cones_oracle.h", "Makefile")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp, u32 = C.c_void_p, C.c_uint32
        L.orc_from_msg.argtypes = [C.POINTER(View), vp]
        L.orc_sector_of.argtypes = [C.c_float, C.c_float]
        L.orc_atan2f.argtypes = [C.c_float, C.c_float]
        L.orc_atan2f.restype = C.c_float
        L.orc_time_atan2f.argtypes = [vp, vp, u32, C.c_int, vp]
        L.orc_time_atan2f.restype = C.c_double
        L.orc_ground_minima.argtypes = [vp, u32, C.c_float, vp]
        L.orc_ground_minima.restype = None
        L.orc_ground_mask.argtypes = [vp, u32, vp, vp]
        L.orc_ground_mask.restype = None
        L.orc_ground_node.argtypes = [C.POINTER(View), C.POINTER(GroundParams), vp, C.POINTER(u32), vp, vp]
        L.orc_crop_mask.argtypes = [vp, u32, C.POINTER(DetectParams), vp]
        L.orc_crop_mask.restype = None
        L.orc_voxel_grid.argtypes = [vp, u32, C.POINTER(DetectParams), C.c_int, vp, vp, vp, C.POINTER(u32),
                                     C.POINTER(Counters)]
        L.orc_extract_clusters.argtypes = [vp, u32, C.POINTER(DetectParams), C.c_int, vp, vp, u32, C.POINTER(u32),
                                           C.POINTER(u32), vp]
        L.orc_label_bruteforce.argtypes = [vp, u32, C.POINTER(DetectParams), vp]
        L.orc_label_bruteforce.restype = None
        L.orc_r2.argtypes = [C.POINTER(DetectParams)]
        L.orc_r2.restype = C.c_float
        L.orc_detect.argtypes = [C.POINTER(View), C.POINTER(DetectParams), C.POINTER(GroundParams), C.c_int, vp, u32,
                                 C.POINTER(u32), C.POINTER(Counters), C.POINTER(Timing)]
        L.orc_extend.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_double]
        L.orc_extend.restype = None
        L.orc_reconstruct_cone.argtypes = [vp, u32, C.c_float, C.c_float, C.c_float, vp, u32]
        L.orc_reconstruct_cone.restype = u32
        L.orc_to_image.argtypes = [vp, u32, vp]
        L.orc_to_image.restype = u32
        _lib = L
    return _lib


def dparams(p) -> DetectParams:
    """Accepts anything with the reference's parameter names (e.g. cones_perception_b200.DetectParams)."""
    return DetectParams(p.distance_treshold_max, p.distance_treshold_min, p.level_threshold, p.angle_threshold,
                        p.voxel_filter_leaf_size_x, p.voxel_filter_leaf_size_y, p.voxel_filter_leaf_size_z,
                        int(p.min_cluster_size), int(p.max_cluster_size), getattr(p, "CONE_WIDTH", 0.228),
                        getattr(p, "CONE_HEIGHT", 0.325))


def gparams(p) -> GroundParams:
    return GroundParams(int(p.num_of_sectors), float(p.default_lowest_point))


def view_of_xyzi(xyzi: np.ndarray, with_intensity: bool = True) -> View:
    a = xyzi
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"] and a.shape[-1] == 4
    n = a.size // 4
    return View(a.ctypes.data, n, 1, 16, 16 * n, 0, 4, 8, 12 if with_intensity else -1, 0, 1)


def view_of_msg(msg, fake_missing_intensity: bool) -> View:
    oi = msg.offset_of("intensity")
    if oi < 0 and fake_missing_intensity:
        oi = 0
    return View(msg.data.ctypes.data, msg.width, msg.height, msg.point_step, msg.row_step, msg.offset_of("x"),
                msg.offset_of("y"), msg.offset_of("z"), oi, 1 if msg.is_bigendian else 0, 1 if msg.is_dense else 0)


def from_msg(view: View) -> np.ndarray:
    out = np.zeros(view.width * view.height, dtype=POINT_DTYPE)
    rc = lib().orc_from_msg(C.byref(view), out.ctypes.data)
    if rc:
        raise ValueError(f"orc_from_msg failed: {rc}")
    return out


def points32(xyzi: np.ndarray) -> np.ndarray:
    """float32 [N,4] -> PCL PointXYZI records."""
    a = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
    out = np.zeros(len(a), dtype=POINT_DTYPE)
    out["x"], out["y"], out["z"], out["intensity"] = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    out["pad"] = 1.0
    return out


def ground_minima(pts: np.ndarray, default_lowest: float) -> np.ndarray:
    low = np.zeros(NSECT, np.float32)
    lib().orc_ground_minima(pts.ctypes.data, len(pts), default_lowest, low.ctypes.data)
    return low


def ground_mask(pts: np.ndarray, low: np.ndarray) -> np.ndarray:
    keep = np.zeros(len(pts), np.uint8)
    low = np.ascontiguousarray(low, np.float32)
    lib().orc_ground_mask(pts.ctypes.data, len(pts), low.ctypes.data, keep.ctypes.data)
    return keep


def ground_node(view: View, g):
    n = view.width * view.height
    out = np.zeros(n, dtype=POINT_DTYPE)
    low = np.zeros(NSECT, np.float32)
    keep = np.zeros(n, np.uint8)
    kept = C.c_uint32()
    gp = gparams(g)
    rc = lib().orc_ground_node(C.byref(view), C.byref(gp), out.ctypes.data, C.byref(kept), low.ctypes.data,
                               keep.ctypes.data)
    if rc:
        raise ValueError(f"orc_ground_node failed: {rc}")
    return out, kept.value, low, keep


def crop_mask(pts: np.ndarray, d) -> np.ndarray:
    keep = np.zeros(len(pts), np.uint8)
    dp = dparams(d)
    lib().orc_crop_mask(pts.ctypes.data, len(pts), C.byref(dp), keep.ctypes.data)
    return keep


def voxel_grid(pts: np.ndarray, d, mode: int = CANONICAL):
    n = len(pts)
    keys = np.zeros(n, np.uint32)
    order = np.zeros(n, np.uint32)
    vox = np.zeros(max(n, 1), dtype=POINT_DTYPE)
    nv = C.c_uint32()
    ctr = Counters()
    dp = dparams(d)
    lib().orc_voxel_grid(pts.ctypes.data, n, C.byref(dp), mode, keys.ctypes.data, order.ctypes.data, vox.ctypes.data,
                         C.byref(nv), C.byref(ctr))
    return keys, order, vox[:nv.value].copy(), ctr


def extract_clusters(vox: np.ndarray, d, mode: int = CANONICAL):
    n = len(vox)
    labels = np.zeros(max(n, 1), np.int32)
    clusters = np.zeros(max(n, 1), dtype=CLUSTER_DTYPE)
    members = np.zeros(max(n, 1), np.uint32)
    k, comps = C.c_uint32(), C.c_uint32()
    dp = dparams(d)
    rc = lib().orc_extract_clusters(vox.ctypes.data, n, C.byref(dp), mode, labels.ctypes.data, clusters.ctypes.data,
                                    max(n, 1), C.byref(k), C.byref(comps), members.ctypes.data)
    if rc:
        raise ValueError(f"orc_extract_clusters failed: {rc}")
    cl = clusters[:k.value].copy()
    return labels[:n].copy(), cl, comps.value, members[:int(cl["size"].sum())].copy()


def label_bruteforce(vox: np.ndarray, d) -> np.ndarray:
    labels = np.zeros(max(len(vox), 1), np.int32)
    dp = dparams(d)
    lib().orc_label_bruteforce(vox.ctypes.data, len(vox), C.byref(dp), labels.ctypes.data)
    return labels[:len(vox)].copy()


def r2(d) -> float:
    dp = dparams(d)
    return float(lib().orc_r2(C.byref(dp)))


def detect(view: View, d, g=None, mode: int = CANONICAL, cap: int = 1 << 16):
    """Returns (clusters, counters, timing)."""
    out = np.zeros(cap, dtype=CLUSTER_DTYPE)
    k = C.c_uint32()
    ctr, tm = Counters(), Timing()
    dp = dparams(d)
    gp = gparams(g) if g is not None else None
    rc = lib().orc_detect(C.byref(view), C.byref(dp), C.byref(gp) if gp is not None else None, mode, out.ctypes.data,
                          cap, C.byref(k), C.byref(ctr), C.byref(tm))
    if rc:
        raise ValueError(f"orc_detect failed: {rc}")
    return out[:k.value].copy(), ctr, tm


def time_atan2f(y: np.ndarray, x: np.ndarray, use_libm: bool) -> float:
    """Seconds for one pass of atan2f over the pairs (restated routine or this box's libm)."""
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    chk = C.c_float()
    return float(lib().orc_time_atan2f(y.ctypes.data, x.ctypes.data, len(y), 1 if use_libm else 0, C.byref(chk)))


def extend(x: float, y: float, length: float):
    cx, cy = C.c_float(x), C.c_float(y)
    lib().orc_extend(C.byref(cx), C.byref(cy), length)
    return cx.value, cy.value


IMG_ROWS, IMG_COLS = 15, 12
CONE_EMPTY, CONE_BAD_INDEX, CONE_BAD_INTENSITY = 1, 2, 4


def reconstruct_cone(pts: np.ndarray, cx: float, cy: float, cone_width: float = 0.228) -> np.ndarray:
    """get_reconstructed_cone (src/cone_detection.cpp:222-238) on PointXYZI records; returns the crop."""
    out = np.zeros(max(len(pts), 1), dtype=POINT_DTYPE)
    m = lib().orc_reconstruct_cone(pts.ctypes.data, len(pts), cx, cy, cone_width, out.ctypes.data, len(out))
    return out[:m].copy()


def to_image(xyzi: np.ndarray):
    """ColorClassifier.to_image (scripts/color_classifier_server.py:130-156); returns (image[15,12] u8, flags)."""
    a = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
    img = np.zeros((IMG_ROWS, IMG_COLS), np.uint8)
    flags = lib().orc_to_image(a.ctypes.data, len(a), img.ctypes.data)
    return img, int(flags)
