"""ctypes binding of oracle/_ref/libconesref.so — the reference's OWN node sources
(/root/reference/src/ground_removal.cpp, cone_detection.cpp, perception_handling/utils.cpp) compiled
unmodified against the stand-in ROS/PCL surface in oracle/ref_shim/.  TEST INFRASTRUCTURE ONLY.

The library can only be (re)built where /root/reference exists (the build container); the GPU box uses the
prebuilt file that travels with the snapshot.  Nothing here reads /root/reference at run time."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libconesref.so")
REFERENCE = "/root/reference"
_lib = None


def available() -> bool:
    return os.path.exists(LIB) or os.path.isdir(os.path.join(REFERENCE, "src"))


def build(force: bool = False) -> str | None:
    """make -C oracle ref; returns the library path, or None when the reference sources are not here."""
    if os.path.isdir(os.path.join(REFERENCE, "src")):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []) + ["ref"], check=True)
    return LIB if os.path.exists(LIB) else None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if build() is None:
            raise FileNotFoundError(f"{LIB} is missing and {REFERENCE} is not present to build it from")
        L = C.CDLL(LIB)
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int32
        L.ref_euclidan_dist.argtypes = [C.c_float] * 6
        L.ref_euclidan_dist.restype = C.c_float
        L.ref_ground_create.argtypes = [C.c_char_p]
        L.ref_ground_create.restype = vp
        L.ref_ground_destroy.argtypes = [vp]
        L.ref_ground_handle.argtypes = [vp, vp, u32, u32, u32, u32, i32, i32, i32, i32, vp, vp, vp, vp]
        L.ref_ground_handle.restype = C.c_int64
        L.ref_detect_create.argtypes = [C.c_char_p, C.c_int]
        L.ref_detect_create.restype = vp
        L.ref_detect_destroy.argtypes = [vp]
        L.ref_detect_handle.argtypes = [vp, vp, u32, u32, u32, u32, i32, i32, i32, i32, vp, vp, u32, vp, vp]
        _lib = L
    return _lib


def _params(d: dict) -> bytes:
    def fmt(v):
        if isinstance(v, bool):
            return "true" if v else "false"
        return repr(float(v)) if isinstance(v, float) else str(v)
    return ";".join(f"~{k}={fmt(v)}" for k, v in d.items()).encode()


def euclidan_dist(a, b) -> np.float32:
    return np.float32(lib().ref_euclidan_dist(*[float(np.float32(v)) for v in (*a, *b)]))


class GroundNode:
    """The real GroundRemover node (src/ground_removal.cpp) behind the in-process message pump."""

    def __init__(self, **params):
        self._h = lib().ref_ground_create(_params(params))

    def handle(self, xyzi: np.ndarray, with_intensity_field: bool = True):
        """One cloud_handler callback on a compact x,y,z,intensity cloud; returns the published [N,8] float32
        PCL-layout cloud plus (point_step, n_fields, stamp_nsec) of the published message."""
        a = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
        n = len(a)
        out = np.zeros((n, 8), np.float32)
        step, nf, nsec = C.c_uint32(), C.c_uint32(), C.c_uint32()
        got = lib().ref_ground_handle(self._h, a.ctypes.data, n, 1, 16, 16 * n, 0, 4, 8, 12 if with_intensity_field else -1,
                                      out.ctypes.data, C.byref(step), C.byref(nf), C.byref(nsec))
        assert got == n, got
        return out, (step.value, nf.value, nsec.value)

    def close(self):
        if self._h:
            lib().ref_ground_destroy(self._h)
            self._h = None


class DetectNode:
    """The real ConeDetector node (src/cone_detection.cpp); VoxelGrid / EuclideanClusterExtraction inside it are
    the oracle's pcl_faithful restatement (PCL is not installed), everything else is the reference's code."""

    def __init__(self, service: bool = True, **params):
        self._h = lib().ref_detect_create(_params(params), 0 if service else -1)

    def handle(self, xyzi: np.ndarray, with_intensity_field: bool = True, cap: int = 4096):
        """One callback; returns the four published clouds as [k,2] float32 arrays (x, y per cone)."""
        a = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
        n = len(a)
        out = np.zeros((4, cap, 2), np.float32)
        counts = np.zeros(4, np.uint32)
        step, nf = C.c_uint32(), C.c_uint32()
        rc = lib().ref_detect_handle(self._h, a.ctypes.data, n, 1, 16, 16 * n, 0, 4, 8, 12 if with_intensity_field else -1,
                                     out.ctypes.data, counts.ctypes.data, cap, C.byref(step), C.byref(nf))
        assert rc == 0, rc
        return [out[k, :counts[k]].copy() for k in range(4)]

    def close(self):
        if self._h:
            lib().ref_detect_destroy(self._h)
            self._h = None
