"""Shared helpers for the parity tests: run the oracle stage by stage and compare with the
taps of libconesgpu."""
from __future__ import annotations

import numpy as np

from cones_perception_b200 import api
from cones_perception_b200.pointcloud2 import PointCloud2
from oracle import oracle as O


def oracle_stages(xyzi: np.ndarray, d, g=None, mode=O.CANONICAL):
    """Reference pipeline on one frame, every intermediate kept (canonical mode by default)."""
    pts = O.points32(xyzi)
    n = len(pts)
    out = {"n": n}
    if g is not None:
        low = O.ground_minima(pts, g.default_lowest_point)
        gkeep = O.ground_mask(pts, low)
        out["low"] = low
        out["ground_keep"] = gkeep
        # the ground node pads back to N with zero points (src/ground_removal.cpp:79)
        kept = pts[gkeep.astype(bool)]
        padded = np.zeros(n, dtype=O.POINT_DTYPE)
        padded["pad"] = 1.0
        padded[:len(kept)] = kept
        src_index = np.concatenate([np.nonzero(gkeep)[0], np.full(n - len(kept), -1)])
        stage_in = padded
    else:
        src_index = np.arange(n)
        stage_in = pts
    ckeep = O.crop_mask(stage_in, d).astype(bool)
    cropped = stage_in[ckeep]
    out["crop_index"] = src_index[ckeep]          # original input index (-1 for filler zeros)
    out["cropped"] = cropped
    keys, order, vox, ctr = O.voxel_grid(cropped, d, mode)
    out.update(keys=keys, order=order, vox=vox, ctr=ctr)
    labels, clusters, comps, members = O.extract_clusters(vox, d, mode)
    out.update(labels=labels, clusters=clusters, n_components=comps, members=members)
    return out


def vox_xyzi(vox: np.ndarray) -> np.ndarray:
    return np.stack([vox["x"], vox["y"], vox["z"], vox["intensity"]], 1)


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_frame_parity(gpu: "api.ConesGpu", f: int, ora: dict, offs: dict, taps: dict, ctr, k_off, clusters):
    """Bit-exact comparison of frame f of a batch run against its oracle stages."""
    c0, c1 = offs["c_off"][f], offs["c_off"][f + 1]
    v0, v1 = offs["v_off"][f], offs["v_off"][f + 1]
    exp_idx = ora["crop_index"]
    got_idx = taps["crop_index"][c0:c1].astype(np.int64)
    got_idx[got_idx == 0xFFFFFFFF] = -1
    # filler zeros: the GPU keeps ONE record with multiplicity; the oracle materialises them
    n_fill = int((exp_idx < 0).sum())
    exp_real = exp_idx[exp_idx >= 0]
    got_real = got_idx[got_idx >= 0]
    assert np.array_equal(got_real, exp_real), f"frame {f}: crop keep-mask differs"
    assert (got_idx < 0).sum() == (1 if n_fill else 0)
    assert ctr["n_cropped"][f] == len(got_idx)
    # G: survivors of the ground node (= N without ground removal), counted although all-ground rows are skipped
    exp_g = int(ora["ground_keep"].sum()) if "ground_keep" in ora else ora["n"]
    assert int(ctr["n_ground_kept"][f]) == exp_g, f"frame {f}: n_ground_kept {ctr['n_ground_kept'][f]} vs {exp_g}"
    if n_fill == 0:
        assert np.array_equal(taps["voxel_keys"][c0:c1], ora["keys"]), f"frame {f}: voxel keys differ"
        assert np.array_equal(taps["voxel_order"][c0:c1] - c0, ora["order"]), f"frame {f}: voxel order differs"
    else:
        assert np.array_equal(np.unique(taps["voxel_keys"][c0:c1]), np.unique(ora["keys"]))
    exp_vox = vox_xyzi(ora["vox"])
    got_vox = taps["voxel_cloud"][v0:v1]
    assert got_vox.shape == exp_vox.shape, f"frame {f}: V {got_vox.shape} vs {exp_vox.shape}"
    assert np.array_equal(bits(got_vox), bits(exp_vox)), f"frame {f}: voxel centroids not bit-exact"
    assert np.array_equal(taps["labels"][v0:v1], ora["labels"]), f"frame {f}: cluster membership differs"
    assert ctr["n_voxels"][f] == len(exp_vox)
    assert ctr["n_components"][f] == ora["n_components"]
    got_cl = clusters[k_off[f]:k_off[f + 1]]
    exp_cl = ora["clusters"]
    assert len(got_cl) == len(exp_cl), f"frame {f}: K {len(got_cl)} vs {len(exp_cl)}"
    assert np.array_equal(got_cl["size"], exp_cl["size"])
    assert np.array_equal(got_cl["min_index"], exp_cl["min_index"])
    assert np.array_equal(bits(got_cl["x"]), bits(exp_cl["x"])), f"frame {f}: cluster x not bit-exact"
    assert np.array_equal(bits(got_cl["y"]), bits(exp_cl["y"])), f"frame {f}: cluster y not bit-exact"


def run_batch_with_taps(gpu: "api.ConesGpu", frames: list, d, g):
    msgs = [PointCloud2.from_xyzi(a) for a in frames]
    gpu.set_host_input(msgs)
    gpu.run(d, g)
    ctr, k_off, clusters = gpu.results()
    offs = {"c_off": gpu.tap(api.TAP_CROP_OFFSETS), "v_off": gpu.tap(api.TAP_VOXEL_OFFSETS)}
    taps = {"crop_index": gpu.tap(api.TAP_CROP_INDEX), "voxel_keys": gpu.tap(api.TAP_VOXEL_KEYS),
            "voxel_order": gpu.tap(api.TAP_VOXEL_ORDER), "voxel_cloud": gpu.tap(api.TAP_VOXEL_CLOUD),
            "labels": gpu.tap(api.TAP_LABELS)}
    if g is not None:
        taps["low"] = gpu.tap(api.TAP_SECTOR_LOW).reshape(-1, 17)
    return ctr, k_off, clusters, offs, taps


def boundary_cloud(d, seed: int = 0, n_random: int = 20000) -> np.ndarray:
    """Points placed on and within a few ulp of every threshold of the path: sector boundaries,
    +-angle_threshold, the +-x / +-y axes, distance_treshold_min/max, level_threshold, plus signed
    zeros, denormals, huge and non-finite values.  Used to prove the guard-band fallbacks exact."""
    rng = np.random.default_rng(seed)
    rows = []
    sa = np.float32((360 // 16) * np.pi / 180)
    theta = d.angle_threshold * np.pi / 180
    angles = [k * float(sa) for k in range(0, 18)] + [theta, -theta, np.pi, -np.pi, np.pi / 2, -np.pi / 2, 0.0,
                                                        2 * np.pi - 1e-7]
    for a in angles:
        for eps in (0.0, 1e-8, -1e-8, 1e-7, -1e-7, 5e-7, -5e-7, 2e-6, -2e-6, 9e-6, -9e-6, 3e-5, -3e-5):
            for r in (0.9, 1.7, 3.3, 6.9, 9.7):
                z = rng.choice([-0.6, -0.601, -0.5, -0.45, 0.0, 0.3])
                rows.append([r * np.cos(a + eps), r * np.sin(a + eps), z])
    # exact axis points and signed zeros
    for x, y in [(1, 0.0), (1, -0.0), (-1, 0.0), (-1, -0.0), (0.0, 1), (-0.0, 1), (0.0, -1), (-0.0, -1), (0.0, 0.0),
                 (-0.0, 0.0), (0.0, -0.0), (-0.0, -0.0), (1e-40, 1e-40), (-1e-40, 2e-39), (3, 1e-42), (3, -1e-42),
                 (1e-20, 5), (-1e-20, 5), (2, 1e-9), (2, -1e-9), (-2, 1e-9), (-2, -1e-9)]:
        for z in (-0.7, -0.2, 0.1):
            rows.append([x * 2.5, y * 2.5, z])
    # distances within a few ulp of dmin / dmax along assorted directions
    for dist in (d.distance_treshold_min, d.distance_treshold_max):
        for _ in range(400):
            v = rng.normal(size=3)
            v[2] = abs(v[2]) * 0.2
            v /= np.linalg.norm(v)
            for k in (-3, -2, -1, 0, 1, 2, 3):
                rows.append(list(v * dist * (1.0 + k * 2.0 ** -24)))
    # level threshold +- ulps
    lv = np.float32(d.level_threshold)
    for zz in (lv, np.nextafter(lv, np.float32(-100)), np.nextafter(lv, np.float32(100))):
        for _ in range(20):
            rows.append([rng.uniform(1.5, 5), rng.uniform(-2, 2), zz])
    # extreme magnitudes and non-finite values
    rows += [[1e30, 1.0, 0.0], [3e38, 3e38, 0.0], [1e-30, 1e-30, 0.1], [np.inf, 1, 0], [1, -np.inf, 0], [1, 1, np.inf],
             [np.nan, 1, 0], [1, np.nan, 0], [1, 1, np.nan], [1, 1, -np.inf], [-np.nan, 2, -0.7]]
    a = np.array(rows, np.float64)
    # random background incl. a ground sheet so every sector has a low minimum
    bg = np.stack([rng.uniform(-9, 9, n_random), rng.uniform(-9, 9, n_random),
                   np.where(rng.random(n_random) < 0.7, -0.6 + rng.normal(0, 0.002, n_random),
                            rng.uniform(-0.55, 1.0, n_random))], 1)
    a = np.concatenate([a, bg])
    rng.shuffle(a)
    out = np.zeros((len(a), 4), np.float32)
    out[:, :3] = a.astype(np.float32)
    out[:, 3] = rng.uniform(0, 100, len(a))
    return out


# ---- the stateful tail of ConeDetector::get_centroid_clouds restated in Python (test oracle) ----
class TrackerReference:
    """src/cone_detection.cpp:276-340 (radial extension, temporal gate / points buffer, colour
    routing) on numpy float32 scalars with the reference's float/double promotion rules."""

    def __init__(self, classify_colors=False, use_points_buffer=False, match=0.5, ext=0.05, forced_color=0,
                 color_fn=None):
        self.classify_colors, self.use_points_buffer = classify_colors, use_points_buffer
        self.match, self.ext, self.forced_color = match, ext, forced_color
        # color_fn(centres) -> colours, possibly FEWER than centres (the service skips empty crops,
        # scripts/color_classifier_server.py:83-84); std::transform then fills the front only (:352-353)
        self.color_fn = color_fn
        self.prev = None                      # prev_detected_cones (NULL before the first frame)
        self.prev_col = [None] * 4            # prev_centroid_clouds

    @staticmethod
    def dist(a, b):                           # perception_handling::euclidan_dist, utils.cpp:32-34
        f32, f64 = np.float32, np.float64
        s = f64(f32(a[0]) - f32(b[0])) ** 2 + f64(f32(a[1]) - f32(b[1])) ** 2 + f64(f32(a[2]) - f32(b[2])) ** 2
        return f32(np.sqrt(s))

    def update(self, centroids):
        f32, f64 = np.float32, np.float64
        clouds = [[] for _ in range(4)]
        current, need = [], []
        for cx, cy in centroids:
            px, py, pz = f32(cx), f32(cy), f32(0.0)
            ln = self.dist((px, py, pz), (0, 0, 0))                                    # :276
            px, py = f32(f64(px) + f64(f32(px / ln)) * self.ext), f32(f64(py) + f64(f32(py / ln)) * self.ext)
            p = (px, py, pz)
            current.append(p)
            if self.prev is not None:                                                  # :282
                for q in self.prev:
                    if (not self.use_points_buffer) or f64(self.dist(p, q)) < self.match:   # :286
                        if self.classify_colors:
                            need_color = True
                            for i in range(1, 4):                                      # :291-306
                                if self.prev_col[i] is not None:
                                    for c in self.prev_col[i]:
                                        if f64(self.dist(p, c)) < self.match:
                                            need_color = False
                                            clouds[i].append(p)
                                            break
                                    if not need_color:
                                        break
                            if need_color:
                                need.append(p)
                        else:
                            clouds[0].append(p)                                        # :315
                        break                                                          # :317
        if self.classify_colors:                                                       # :326-333
            colors = [self.forced_color if self.color_fn is None else 0] * len(need)
            if self.color_fn is not None and need:
                got = self.color_fn(need)
                colors[:len(got)] = got[:len(need)]
            for p, c in zip(need, colors):
                clouds[c].append(p)
        self.prev_col = [list(c) for c in clouds]                                      # :335-337
        self.prev = current                                                            # :339
        return clouds
