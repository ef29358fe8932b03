"""Synthetic LiDAR scans for BASELINE.json's five configs.

Ray-cast scenes come from the multi-threaded C++ generator (scangen/scan_gen.cpp); the
adversarial items of config 5 that are not ray-castable (serpentine chain, solid blobs,
exact-size components) are appended as direct points built with numpy.
Every cloud is float32 [N, 4] = x, y, z, intensity (the compact device layout).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from .params import PRESETS, DetectParams, GroundParams

_HERE = os.path.dirname(os.path.abspath(__file__))
SCAN_LIB = os.path.join(_HERE, "scangen", "libconesscan.so")


class _Sensor(C.Structure):
    _fields_ = [("beams", C.c_uint32), ("az", C.c_uint32), ("sweeps", C.c_uint32),
                ("elev_min_deg", C.c_float), ("elev_max_deg", C.c_float), ("sensor_h", C.c_float),
                ("max_range", C.c_float), ("noise_sigma", C.c_float)]


class _Scene(C.Structure):
    _fields_ = [("cones", C.c_void_p), ("ncones", C.c_uint32), ("walls", C.c_void_p), ("nwalls", C.c_uint32),
                ("posts", C.c_void_p), ("nposts", C.c_uint32)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(SCAN_LIB):
            from .build import build_scangen
            build_scangen()
        _lib = C.CDLL(SCAN_LIB)
        _lib.scan_generate_batch.argtypes = [C.POINTER(_Sensor), C.POINTER(_Scene), C.c_uint64, C.c_uint32, C.c_int,
                                             C.c_void_p, C.c_int]
        _lib.scan_points_per_frame.argtypes = [C.POINTER(_Sensor)]
        _lib.scan_points_per_frame.restype = C.c_uint64
    return _lib


@dataclass
class ScanConfig:
    name: str
    beams: int
    az: int
    elev_min_deg: float
    elev_max_deg: float
    sensor_h: float
    preset: str
    ground_removal: bool
    sweeps: int = 1
    max_range: float = 60.0
    noise_sigma: float = 0.005
    jitter: bool = False
    cones: np.ndarray = field(default_factory=lambda: track_cones())
    walls: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.float32))
    posts: np.ndarray = field(default_factory=lambda: np.zeros((0, 5), np.float32))

    @property
    def points_per_frame(self) -> int:
        return self.beams * self.az * self.sweeps

    @property
    def detect(self) -> DetectParams:
        return PRESETS[self.preset]

    @property
    def ground(self) -> GroundParams | None:
        return GroundParams() if self.ground_removal else None


def track_cones(n_per_row: int = 20, spacing: float = 1.0, half_width: float = 1.5, x0: float = 1.5) -> np.ndarray:
    """40 cones in two rows 3 m apart, ahead of the sensor (an FS track segment)."""
    xs = x0 + spacing * np.arange(n_per_row, dtype=np.float32)
    left = np.stack([xs, np.full_like(xs, half_width)], 1)
    right = np.stack([xs + 0.5 * spacing, np.full_like(xs, -half_width)], 1)
    return np.concatenate([left, right]).astype(np.float32)


def config(idx: int) -> ScanConfig:
    """BASELINE.json configs[idx-1]."""
    if idx == 1:   # 16 x 1875 = 30 000 rays, params "our", no ground removal
        return ScanConfig("cfg1_fs_track_30k", 16, 1875, -15.0, 15.0, 0.6, "our", False)
    if idx in (2, 3):  # 64 x 2048 = 131 072 rays, params "simulation", ground removal on
        return ScanConfig("cfg2_fs_track_130k" if idx == 2 else "cfg3_batch_130k", 64, 2048, -24.8, 2.0, 0.6,
                          "simulation", True, jitter=(idx == 3))
    if idx == 4:   # 128 x 2048 x 10 sweeps = 2.62 M rays, params "fsai", walls inside 6 m
        walls = np.array([[1.0, 3.5, 5.5, 3.5, -0.15, 1.2], [1.0, -3.5, 5.5, -3.5, -0.15, 1.2],
                          [5.5, -3.5, 5.5, -1.0, -0.15, 0.8], [5.5, 1.0, 5.5, 3.5, -0.15, 0.8]], np.float32)
        return ScanConfig("cfg4_dense_2p6M", 128, 2048, -25.0, 15.0, 0.15, "fsai", False, sweeps=10, walls=walls)
    if idx == 5:   # cfg 2 scene + adversarial items (direct points added by adversarial_points())
        posts = np.array([[2.0 + 0.35 * i, 4.0, 0.03, -0.6, 0.4] for i in range(86)] +
                         [[2.0 + 0.35 * i, -4.2, 0.03, -0.6, 0.4] for i in range(86)], np.float32)
        return ScanConfig("cfg5_adversarial", 64, 2048, -24.8, 2.0, 0.6, "simulation", True, posts=posts)
    raise ValueError("config index must be 1..5")


def generate(cfg: ScanConfig, frames: int = 1, base_seed: int = 0, nthreads: int | None = None,
             out: np.ndarray | None = None) -> np.ndarray:
    """Returns float32 [frames, N, 4]; frame f uses seed base_seed + f."""
    lib = _load()
    s = _Sensor(cfg.beams, cfg.az, cfg.sweeps, cfg.elev_min_deg, cfg.elev_max_deg, cfg.sensor_h, cfg.max_range,
                cfg.noise_sigma)
    cones = np.ascontiguousarray(cfg.cones, np.float32)
    walls = np.ascontiguousarray(cfg.walls, np.float32)
    posts = np.ascontiguousarray(cfg.posts, np.float32)
    sc = _Scene(cones.ctypes.data, len(cones), walls.ctypes.data, len(walls), posts.ctypes.data, len(posts))
    n = cfg.points_per_frame
    if out is None:
        out = np.empty((frames, n, 4), dtype=np.float32)
    assert out.dtype == np.float32 and out.size == frames * n * 4 and out.flags["C_CONTIGUOUS"]
    if nthreads is None:
        nthreads = min(os.cpu_count() or 1, 32)
    rc = lib.scan_generate_batch(C.byref(s), C.byref(sc), base_seed, frames, 1 if cfg.jitter else 0, out.ctypes.data,
                                 nthreads)
    if rc != 0:
        raise RuntimeError("scan_generate_batch failed")
    return out.reshape(frames, n, 4)


# ---- config 5: direct (non ray-cast) adversarial structures ----------------------------------
def _voxel_centres(i, j, k, leaf=0.04):
    return np.stack([(i + 0.5) * leaf, (j + 0.5) * leaf, (k + 0.5) * leaf], -1).astype(np.float32)


def adversarial_points(d: DetectParams, seed: int = 0) -> np.ndarray:
    """Serpentine chain (deep union-find), two solid blobs (> max, ~2000 neighbours/voxel),
    components of exactly max_cluster_size and max_cluster_size + 1 voxels."""
    rng = np.random.default_rng(seed)
    parts = []
    # (ii) serpentine: rows along y, 0.06 m steps (< tol 0.397), rows 0.48 m apart (> tol), joined at alternating ends
    step, gap = 0.06, 0.48
    ys = np.arange(-4.5, 4.5 + 1e-6, step)
    z_layers = np.arange(1.2, 4.4, gap)
    xs = np.arange(1.8, 6.4, gap)
    chain = []
    flip = False
    for zi, z in enumerate(z_layers):
        xs_l = xs if zi % 2 == 0 else xs[::-1]
        for xi, x in enumerate(xs_l):
            row = ys[::-1] if flip else ys
            if xi > 0:  # bridge point between consecutive rows (rows are > tol apart)
                chain.append((0.5 * (x + xs_l[xi - 1]), row[0], z))
            chain.extend((x, y, z) for y in row)
            flip = not flip
        # bridge to the next layer at the last (x, y)
        if zi + 1 < len(z_layers):
            x_last, y_last = chain[-1][0], chain[-1][1]
            chain.append((x_last, y_last, z + gap / 2))
    parts.append(np.array(chain, np.float32))
    # (iii) two solid 1 m^3 blobs on the voxel lattice (25^3 voxels each)
    g = np.arange(25)
    gi, gj, gk = np.meshgrid(g, g, g, indexing="ij")
    for (bx, by, bz) in ((50, 150, -12), (120, -180, 30)):
        parts.append(_voxel_centres(gi.ravel() + bx, gj.ravel() + by, gk.ravel() + bz))
    # (iv) flat patches of exactly max and max+1 voxels (one point per voxel)
    for extra, (bx, by) in ((0, (60, 60)), (1, (60, -90))):
        n = d.max_cluster_size + extra
        idx = np.arange(n)
        parts.append(_voxel_centres(bx + idx % 25, by + idx // 25, np.full(n, 10)))
    pts = np.concatenate(parts)
    inten = rng.uniform(0, 100, len(pts)).astype(np.float32)
    return np.concatenate([pts, inten[:, None]], 1).astype(np.float32)


def generate_config5(frames: int = 1, base_seed: int = 0) -> np.ndarray:
    cfg = config(5)
    base = generate(cfg, frames, base_seed)
    extra = adversarial_points(cfg.detect, base_seed)
    # interleave the direct points with the scan so they are not one contiguous block
    out = []
    for f in range(frames):
        rng = np.random.default_rng(base_seed + f + 12345)
        pos = np.sort(rng.integers(0, base.shape[1], len(extra)))
        merged = np.insert(base[f], pos, extra, axis=0)
        out.append(merged)
    return np.stack(out).astype(np.float32)
