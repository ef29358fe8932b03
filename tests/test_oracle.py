"""CPU tests of the oracle: known-answer tests derived from SURVEY.md Appendix A, cross-checks
between its two modes and against independent implementations, and the committed goldens."""
import hashlib
import os

import numpy as np
import pytest

from cones_perception_b200 import scans
from cones_perception_b200.params import PRESETS, DetectParams, GroundParams
from oracle import oracle as O
from tests.util import oracle_stages, vox_xyzi

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def f32(x):
    return np.float32(x)


def bits(x):
    return int(np.array(x, np.float32).view(np.uint32))


def cloud(rows):
    a = np.zeros((len(rows), 4), np.float32)
    a[:, :len(rows[0])] = np.array(rows, np.float32)
    return a


# ------------------------------------------------------------------ constants (Appendix A table)
def test_constants_bit_patterns():
    assert bits(0.325) == 0x3EA66666 and bits(0.228) == 0x3E6978D5
    tol = np.sqrt(np.float64(np.float32(0.325)) ** 2 + np.float64(np.float32(0.228)) ** 2)
    assert bits(tol) == 0x3ECB4395
    assert bits(O.r2(PRESETS["our"])) == 0x3E216440
    assert bits(np.float32(1.0) / np.float32(0.04)) == 0x41C80000          # inverse leaf = 25.0f exactly
    assert bits((360 // 16) * np.pi / 180) == 0x3EC49809                    # 22 deg, not 22.5


def test_sector_index_range_and_wrap():
    L = O.lib()
    assert L.orc_sector_of(1.0, 0.0) == 0
    assert L.orc_sector_of(1.0, -1e-9) == 16          # angles just below 2*pi land in the 17th sector (Q2)
    assert L.orc_sector_of(-1.0, 0.0) == int(np.floor(np.float32(np.pi) / np.float32(0.38397244)))
    ang = np.linspace(-np.pi, np.pi, 20001)
    s = [L.orc_sector_of(float(np.cos(a)), float(np.sin(a))) for a in ang]
    assert min(s) == 0 and max(s) == 16
    assert L.orc_sector_of(0.0, 0.0) == 0             # atan2(0, 0) = 0: the filler point


# ------------------------------------------------------------------ ground removal (A.2)
def test_ground_minima_and_mask():
    g = GroundParams()
    pts = O.points32(cloud([[5, 0.1, -0.60], [5, 0.2, -0.62], [5, 0.3, -0.51], [5, 0.4, -0.53],
                            [-5, 0.1, 0.3], [0, 5, -0.05]]))
    low = O.ground_minima(pts, g.default_lowest_point)
    assert low[0] == f32(-0.62)
    assert low[8] == f32(-0.1)                         # (-5, 0.1): z = 0.3 does not lower the default
    assert low[4] == f32(-0.1)
    keep = O.ground_mask(pts, low)
    # drop iff (double)z < (double)low + 0.1 : low[0]+0.1 = -0.52 -> -0.53 dropped, -0.51 kept
    # (0, 5, -0.05): low[4] stays at the default -0.1, threshold 0.0 -> dropped
    assert keep.tolist() == [0, 0, 1, 0, 1, 0]


def test_ground_node_pads_with_zero_points():
    a = cloud([[5, 0.1, -0.6, 7], [5, 0.2, 0.5, 8], [4, -3, -0.6, 9], [4, -3.1, 0.4, 10]])
    out, kept, low, keep = O.ground_node(O.view_of_xyzi(a), GroundParams())
    assert kept == 2 and keep.tolist() == [0, 1, 0, 1]
    assert out["x"].tolist() == [5, 4, 0, 0] and out["intensity"].tolist() == [8, 10, 0, 0]
    assert out["pad"].tolist() == [1, 1, 1, 1]


def test_all_ground_cloud_yields_nothing():
    rng = np.random.default_rng(0)
    a = np.zeros((2000, 4), np.float32)
    a[:, 0] = rng.uniform(2, 6, 2000)
    a[:, 1] = rng.uniform(-3, 3, 2000)
    a[:, 2] = -0.6 + rng.normal(0, 0.002, 2000)
    cl, ctr, _ = O.detect(O.view_of_xyzi(a), PRESETS["simulation"], GroundParams())
    assert ctr.n_ground_kept == 0 and ctr.n_cropped == 0 and len(cl) == 0


# ------------------------------------------------------------------ crop (A.3)
def test_crop_thresholds():
    d = PRESETS["our"]  # dmax 7, dmin 0.7, level -0.5, angle 90
    rows = [[3, 0, 0], [3, 0, -0.5], [3, 0, np.nextafter(f32(-0.5), f32(-1))],      # level: < is strict
            [7, 0, 0], [np.nextafter(f32(7), f32(8)), 0, 0],                          # d > dmax strict
            [0.7, 0, 0], [np.nextafter(f32(0.7), f32(0)), 0, 0],                      # d < dmin strict
            [1e-9, 3, 0], [0, 3, 0], [-1e-3, 3, 0], [1e-9, -3, 0], [0, -3, 0],        # |angle| < 90 deg
            [np.nan, 1, 0], [1, np.inf, 0]]
    keep = O.crop_mask(O.points32(cloud(rows)), d)
    assert keep.tolist() == [1, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0] or \
        keep.tolist() == [1, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0]
    # (float)0.7 = 0.69999999 < 0.7 (double): the d < dmin test drops it
    assert keep[5] == 0
    # atan2(3, 1e-9) rounds to float(pi/2) >= theta: dropped, like x = 0
    assert keep[7] == 0


def test_zero_point_is_cropped_by_shipped_presets():
    z = O.points32(np.zeros((1, 4), np.float32))
    for d in PRESETS.values():
        assert O.crop_mask(z, d)[0] == 0


# ------------------------------------------------------------------ VoxelGrid (A.4)
def test_voxel_keys_and_centroids():
    d = PRESETS["our"]
    a = cloud([[1.00, 1.00, 0.00, 10], [1.01, 1.01, 0.01, 20], [1.039, 1.0, 0.0, 30],   # same voxel (25,25,0)
               [1.041, 1.0, 0.0, 40],                                                   # next voxel in x
               [1.0, 1.0, 0.041, 50]])                                                  # next voxel in z
    keys, order, vox, ctr = O.voxel_grid(O.points32(a), d)
    assert list(ctr.min_b) == [25, 25, 0] and list(ctr.div_b) == [2, 1, 2]
    assert keys.tolist() == [0, 0, 0, 1, 2] and order.tolist() == [0, 1, 2, 3, 4]
    assert len(vox) == 3
    sx = (f32(1.00) + f32(1.01)) + f32(1.039)
    assert vox["x"][0] == sx / f32(3) and vox["intensity"][0] == f32(20)
    assert vox["x"][1] == f32(1.041) and vox["z"][2] == f32(0.041)


def test_voxel_boundary_one_ulp():
    d = PRESETS["our"]
    edge = f32(0.08)                      # 2 * 0.04: floor(x * 25) switches from 1 to 2 around here
    lo, hi = np.nextafter(edge, f32(0)), np.nextafter(edge, f32(1))
    a = cloud([[0.0, 0, 0], [lo, 0, 0], [edge, 0, 0], [hi, 0, 0]])
    keys, _, _, _ = O.voxel_grid(O.points32(a), d)
    expect = [int(np.floor(f32(v) * f32(25.0))) for v in (0.0, lo, edge, hi)]
    assert keys.tolist() == expect


def test_voxel_empty_and_passthrough():
    d = PRESETS["our"]
    keys, order, vox, ctr = O.voxel_grid(O.points32(np.zeros((0, 4), np.float32)), d)
    assert len(vox) == 0
    # extents so large that dx*dy*dz overflows int32: PCL returns the input unchanged
    a = cloud([[0, 0, 0], [1000, 1000, 1000], [1000, 1000, 1000.01]])
    keys, order, vox, ctr = O.voxel_grid(O.points32(a), d)
    assert ctr.passthrough == 1 and len(vox) == 3 and vox["x"].tolist() == [0, 1000, 1000]


def test_canonical_vs_faithful_voxel_centroids_within_tolerance():
    cfg = scans.config(2)
    frame = scans.generate(cfg, 1, 11)[0]
    a = oracle_stages(frame, cfg.detect, cfg.ground, O.CANONICAL)
    b = oracle_stages(frame, cfg.detect, cfg.ground, O.PCL_FAITHFUL)
    assert np.array_equal(a["keys"], b["keys"])
    assert np.abs(vox_xyzi(a["vox"])[:, :3] - vox_xyzi(b["vox"])[:, :3]).max() <= 1e-5   # north_star tolerance
    assert np.array_equal(a["labels"], b["labels"])                                       # membership exact
    ca = np.sort(a["clusters"], order=["size", "min_index"])
    cb = np.sort(b["clusters"], order=["size", "min_index"])
    assert np.array_equal(ca["min_index"], cb["min_index"]) and np.array_equal(ca["size"], cb["size"])
    assert np.abs(ca["x"] - cb["x"]).max() <= 1e-5 and np.abs(ca["y"] - cb["y"]).max() <= 1e-5


# ------------------------------------------------------------------ clustering (A.5)
def vox_from(rows):
    return O.points32(cloud(rows))


def test_pair_just_inside_and_outside_r2():
    d = PRESETS["simulation"]
    r2 = f32(O.r2(d))
    r = np.sqrt(np.float64(r2))
    inside = np.nextafter(f32(r), f32(0))
    while f32(inside) * f32(inside) >= r2:
        inside = np.nextafter(inside, f32(0))
    outside = f32(r)
    while f32(outside) * f32(outside) < r2:
        outside = np.nextafter(outside, f32(1))
    lab, cl, comps, _ = O.extract_clusters(vox_from([[0, 0, 0], [inside, 0, 0]]), d)
    assert lab.tolist() == [0, 0] and comps == 1 and cl["size"].tolist() == [2]
    lab, cl, comps, _ = O.extract_clusters(vox_from([[0, 0, 0], [outside, 0, 0]]), d)
    assert lab.tolist() == [0, 1] and comps == 2 and len(cl) == 0      # singletons < min_cluster_size 2


def test_chain_spacing():
    d = PRESETS["simulation"]
    k = 40
    near = vox_from([[0.39 * i, 0, 0] for i in range(k)])
    far = vox_from([[0.40 * i, 0, 0] for i in range(k)])
    lab, cl, comps, _ = O.extract_clusters(near, d)
    assert comps == 1 and cl["size"].tolist() == [k] and cl["min_index"].tolist() == [0]
    lab, cl, comps, _ = O.extract_clusters(far, d)
    assert comps == k and len(cl) == 0


def test_max_cluster_size_kept_and_dropped_whole():
    d = DetectParams(**{**PRESETS["our"].__dict__})
    d.max_cluster_size = 50
    line = lambda n, y: [[0.05 * i, y, 0] for i in range(n)]
    vox = vox_from(line(50, 0.0) + line(51, 5.0) + line(2, 10.0) + line(3, 15.0))
    lab, cl, comps, members = O.extract_clusters(vox, d)
    assert comps == 4
    assert cl["size"].tolist() == [50, 3] and cl["min_index"].tolist() == [0, 103]   # 51 dropped whole, 2 < min
    assert members[:50].tolist() == list(range(50))


def test_cluster_order_ties_by_min_index():
    d = PRESETS["our"]
    tri = lambda x: [[x, 0, 0], [x + 0.1, 0, 0], [x, 0.1, 0]]
    vox = vox_from(tri(10) + tri(0) + tri(5) + [[20, 0, 0], [20.1, 0, 0], [20, 0.1, 0], [20.1, 0.1, 0]])
    _, cl, _, _ = O.extract_clusters(vox, d)
    assert cl["size"].tolist() == [4, 3, 3, 3] and cl["min_index"].tolist() == [9, 0, 3, 6]


def test_labels_match_bruteforce_and_scipy():
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree
    cfg = scans.config(2)
    frame = scans.generate(cfg, 1, 21)[0]
    st = oracle_stages(frame, cfg.detect, cfg.ground)
    vox, lab = st["vox"], st["labels"]
    assert np.array_equal(lab, O.label_bruteforce(vox, cfg.detect))
    xyz = np.stack([vox["x"], vox["y"], vox["z"]], 1).astype(np.float64)
    r2 = O.r2(cfg.detect)
    # guard band: no pair within 1e-6 of r2, so float/double disagreement cannot matter
    tree = cKDTree(xyz)
    pairs = tree.query_pairs(np.sqrt(r2) + 1e-3, output_type="ndarray")
    d2 = ((xyz[pairs[:, 0]] - xyz[pairs[:, 1]]) ** 2).sum(1)
    assert np.abs(d2 - r2).min() > 1e-6
    edges = pairs[d2 < r2]
    n = len(vox)
    ncomp, sl = connected_components(coo_matrix((np.ones(len(edges)), (edges[:, 0], edges[:, 1])), shape=(n, n)),
                                     directed=False)
    first = np.full(ncomp, n)
    np.minimum.at(first, sl, np.arange(n))
    assert np.array_equal(first[sl], lab)


def test_faithful_mode_on_adversarial_scene():
    cfg = scans.config(5)
    frame = scans.generate_config5(1, 0)[0]
    a = oracle_stages(frame, cfg.detect, cfg.ground, O.CANONICAL)
    b = oracle_stages(frame, cfg.detect, cfg.ground, O.PCL_FAITHFUL)
    assert np.array_equal(a["labels"], b["labels"])
    assert np.bincount(a["labels"]).max() >= 10_000          # the serpentine chain is one deep component
    assert (a["clusters"]["size"] == cfg.detect.max_cluster_size).sum() == 1
    assert not (a["clusters"]["size"] > cfg.detect.max_cluster_size).any()


# ------------------------------------------------------------------ centroid + extension (A.6, A.7)
def test_cluster_centroid_is_sequential_fp32():
    d = PRESETS["our"]
    rows = [[1.1, 2.1, 0], [1.2, 2.2, 0], [1.3, 2.3, 0.1]]
    _, cl, _, _ = O.extract_clusters(vox_from(rows), d)
    x = (f32(0) + f32(1.1) + f32(1.2)) + f32(1.3)
    assert cl["x"][0] == x / f32(3)


def test_radial_extension():
    x, y = O.extend(3.0, 4.0, 0.05)
    assert x == f32(np.float64(f32(3.0)) + np.float64(f32(3.0) / f32(5.0)) * 0.05)
    assert y == f32(np.float64(f32(4.0)) + np.float64(f32(4.0) / f32(5.0)) * 0.05)


# ------------------------------------------------------------------ fixtures
def test_real_cone_crops_each_form_one_cluster():
    """The 577 hand-labelled cone crops shipped with the reference (cones_clouds/cones.pkl):
    every crop must voxelise and cluster into exactly one component."""
    z = np.load(os.path.join(GOLD, "cone_crops.npz"))
    pts, lens = z["points"], z["lengths"]
    d = PRESETS["fsai"]
    off = 0
    single = 0
    for n in lens:
        crop = pts[off:off + n]
        off += n
        _, _, vox, _ = O.voxel_grid(O.points32(crop), d)
        lab, _, comps, _ = O.extract_clusters(vox, d)
        single += comps == 1
    assert single == len(lens)


@pytest.mark.parametrize("idx,seed", [(1, 0), (2, 0), (2, 7), (4, 0), (5, 0)])
def test_oracle_matches_committed_goldens(idx, seed):
    z = np.load(os.path.join(GOLD, f"cfg{idx}_seed{seed}_golden.npz"))
    cfg = scans.config(idx)
    frame = scans.generate_config5(1, seed)[0] if idx == 5 else scans.generate(cfg, 1, seed)[0]
    assert hashlib.sha256(frame.tobytes()).hexdigest() == str(z["input_sha256"]), "scan generator drifted"
    cl, ctr, _ = O.detect(O.view_of_xyzi(frame), cfg.detect, cfg.ground, O.CANONICAL)
    assert np.array_equal(cl.view(np.uint32), z["clusters"].view(np.uint32))
    got = [ctr.n_points, ctr.n_ground_kept, ctr.n_cropped, ctr.n_voxels, ctr.n_components, ctr.n_clusters,
           ctr.key_bits]
    assert got == z["counters"].tolist()


def test_generator_is_thread_count_independent():
    cfg = scans.config(3)
    a = scans.generate(cfg, 3, 5, nthreads=1)
    b = scans.generate(cfg, 3, 5, nthreads=3)
    assert np.array_equal(a, b)
    assert not np.array_equal(a[0], a[1])


# ------------------------------------------------------------------ independent numpy restatements
def _np_sector(x, y):
    a = np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(np.float32)
    ang = np.where(a < 0, (a.astype(np.float64) + 2 * np.pi).astype(np.float32), a)
    sa = np.float32((360 // 16) * np.pi / 180)
    return np.floor(ang / sa).astype(np.int64)


def test_ground_and_crop_match_vectorised_numpy():
    """The C++ oracle's masks against a vectorised numpy restatement of the same reference lines
    (src/ground_removal.cpp:58-77, src/cone_detection.cpp:189-204) on a full 130k-point scan."""
    cfg = scans.config(2)
    f = scans.generate(cfg, 1, 33)[0]
    x, y, z = f[:, 0], f[:, 1], f[:, 2]
    pts = O.points32(f)
    s = _np_sector(x, y)
    low = np.full(17, np.float32(-0.1))
    np.minimum.at(low, s, z)
    assert np.array_equal(O.ground_minima(pts, -0.1), low)
    keep_g = ~(z.astype(np.float64) < low[s].astype(np.float64) + 0.1)
    assert np.array_equal(O.ground_mask(pts, low).astype(bool), keep_g)
    for d in PRESETS.values():
        dist = np.sqrt(x.astype(np.float64) ** 2 + y.astype(np.float64) ** 2 + z.astype(np.float64) ** 2).astype(np.float32)
        a = np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(np.float32).astype(np.float64)
        th = d.angle_threshold * np.pi / 180
        drop = (z.astype(np.float64) < d.level_threshold) | (dist.astype(np.float64) > d.distance_treshold_max) | \
               (dist.astype(np.float64) < d.distance_treshold_min) | (-th >= a) | (a >= th)
        assert np.array_equal(O.crop_mask(pts, d).astype(bool), ~drop)


def test_voxel_grid_matches_numpy_restatement():
    """pcl::VoxelGrid keys, voxel order and counts against numpy (np.unique on the PCL idx); centroids
    within float rounding of the float64 mean (the oracle sums sequentially in fp32)."""
    cfg = scans.config(2)
    f = scans.generate(cfg, 1, 34)[0]
    st = oracle_stages(f, cfg.detect, cfg.ground)
    c = st["cropped"]
    xyz = np.stack([c["x"], c["y"], c["z"]], 1)
    inv = np.float32(1.0) / np.float32(0.04)
    cell = np.floor(xyz * inv).astype(np.int64)
    min_b = np.floor(xyz.min(0) * inv).astype(np.int64)
    max_b = np.floor(xyz.max(0) * inv).astype(np.int64)
    div = max_b - min_b + 1
    idx = (cell[:, 0] - min_b[0]) + (cell[:, 1] - min_b[1]) * div[0] + (cell[:, 2] - min_b[2]) * div[0] * div[1]
    order = np.argsort(idx, kind="stable")
    assert np.array_equal(st["keys"], idx[order].astype(np.uint32))
    assert np.array_equal(st["order"], order.astype(np.uint32))
    uniq, inverse, counts = np.unique(idx, return_inverse=True, return_counts=True)
    assert len(st["vox"]) == len(uniq)
    mean = np.zeros((len(uniq), 3))
    np.add.at(mean, inverse, xyz.astype(np.float64))
    mean /= counts[:, None]
    assert np.abs(vox_xyzi(st["vox"])[:, :3] - mean).max() < 2e-6


# ---- colour path inputs: PINNED against the reference's own numpy code --------

def _golden_images():
    z = np.load(os.path.join(GOLD, "cone_images.npz"))
    crops = np.load(os.path.join(GOLD, "cone_crops.npz"))
    off = np.concatenate([[0], np.cumsum(crops["lengths"])])
    real = [crops["points"][off[i]:off[i + 1]] for i in range(len(crops["lengths"]))]
    soff = np.concatenate([[0], np.cumsum(z["synth_lengths"])])
    synth = [z["synth_points"][soff[i]:soff[i + 1]] for i in range(len(z["synth_lengths"]))]
    return z, real, synth


def test_to_image_matches_reference_golden():
    """cone_images.npz was produced by the reference's ColorClassifier.to_image itself
    (tests/golden/make_golden.py::cone_images); the restatement must agree byte for byte."""
    z, real, synth = _golden_images()
    for i, c in enumerate(real):
        img, fl = O.to_image(c)
        assert fl == 0 and np.array_equal(img, z["real_images"][i]), i
    n_raised = 0
    for i, c in enumerate(synth):
        img, fl = O.to_image(c)
        raised = int(z["synth_raised"][i])
        if raised:  # numpy IndexError (2) / interp1d ValueError (4): flagged, image cleared
            assert fl & raised and not img.any(), (i, fl, raised)
            n_raised += 1
        else:
            assert fl == 0 and np.array_equal(img, z["synth_images"][i]), i
    assert n_raised >= 16
    # repeated pixels are common in the real crops, so "the last point wins" is exercised
    assert sum(len(c) for c in real) > 1.5 * int(np.count_nonzero(z["real_images"]))


def test_to_image_empty_and_quirks():
    img, fl = O.to_image(np.zeros((0, 4), np.float32))
    assert fl == O.CONE_EMPTY and not img.any()
    # slope_vert is negative: elevation +1 deg -> round(-0.5 * 16) = -8 -> numpy wraps to row 7
    p = np.array([[10.0, 0.0, 10.0 * np.tan(np.radians(1.0)), 42.9]], np.float32)
    img, fl = O.to_image(p)
    assert fl == 0 and img[7, 0] == 42 and np.count_nonzero(img) == 1
    # elevation -15 deg and +15 deg both land on row 0 (index 0 and index -15)
    for elev in (-14.999, 14.999):
        p = np.array([[10.0, 0.0, 10.0 * np.tan(np.radians(elev)), 7.0]], np.float32)
        img, fl = O.to_image(p)
        assert fl == 0 and img[0, 0] == 7


def test_reconstruct_cone_box_is_inclusive_in_double():
    cw = np.float32(0.228)
    hw = float(cw) / 1.5
    cx, cy = np.float32(3.25), np.float32(-1.5)
    hi = np.float32(float(cx) + hw)
    if float(hi) > float(cx) + hw:
        hi = np.nextafter(hi, np.float32(-np.inf))
    above = np.nextafter(hi, np.float32(np.inf))
    lo = np.float32(float(cx) - hw)
    if float(lo) < float(cx) - hw:
        lo = np.nextafter(lo, np.float32(np.inf))
    below = np.nextafter(lo, np.float32(-np.inf))
    xs = np.array([hi, above, lo, below, cx, np.nan], np.float32)
    pts = O.points32(np.stack([xs, np.full(6, cy), np.arange(6), np.arange(6) * 10], 1).astype(np.float32))
    crop = O.reconstruct_cone(pts, float(cx), float(cy), float(cw))
    assert list(crop["z"]) == [0.0, 2.0, 4.0]           # cloud order kept; NaN and the outer neighbours dropped
    assert list(crop["intensity"]) == [0.0, 20.0, 40.0] and np.all(crop["pad"] == 1.0)


def atan2_cases(seed=0, n=400000):
    """Pairs for the atan2f checks: random bit patterns, lidar-like magnitudes, near-axis ratios, specials."""
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 1 << 32, (n, 2), dtype=np.uint64).astype(np.uint32).view(np.float32)
    xy = rng.uniform(-60, 60, (n, 2)).astype(np.float32)
    near = xy.copy()
    near[: n // 2, 0] *= np.float32(1e-7)
    near[n // 2:, 1] *= np.float32(1e-7)
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3e38, -3e38, 1.7, -1.7e-8, 2.4375,
                   0.4375, 0.6875, 1.1875, 2.0 ** 25, 2.0 ** -29], np.float32)
    grid = np.array([(a, b) for a in sp for b in sp], np.float32)
    allp = np.concatenate([bits, xy, near, grid])
    return np.ascontiguousarray(allp[:, 0]), np.ascontiguousarray(allp[:, 1])


def test_atan2f_restatement_matches_libm():
    """The reference's atan2(float, float) is libm atan2f.  The oracle restates fdlibm's routine; on glibc <= 2.40
    (this image: 2.39; Noetic: 2.31) it must be libm's result bit for bit, including where that is not the
    correctly rounded value."""
    import ctypes as C
    import platform
    libm = C.CDLL("libm.so.6")
    libm.atan2f.argtypes = [C.c_float, C.c_float]
    libm.atan2f.restype = C.c_float
    ver = tuple(int(v) for v in platform.libc_ver()[1].split(".")[:2])
    if ver >= (2, 41):
        pytest.skip("glibc >= 2.41 ships a correctly rounded atan2f; the reference's target (2.31) does not")
    y, x = atan2_cases(n=60000)
    f = O.lib().orc_atan2f
    for a, b in zip(y.tolist(), x.tolist()):
        got, exp = np.float32(f(a, b)), np.float32(libm.atan2f(a, b))
        assert got.view(np.uint32) == exp.view(np.uint32) or (np.isnan(got) and np.isnan(exp)), (a, b)
    # the case that separates it from a correctly rounded atan2: one float BELOW pi/2
    assert np.float32(f(1.7, np.float32(-1.7e-8))).view(np.uint32) == 0x3fc90fda
    assert np.float32(np.arctan2(np.float64(np.float32(1.7)), np.float64(np.float32(-1.7e-8)))).view(np.uint32) == 0x3fc90fdb


# ---- goldens produced by the reference's own compiled node sources (tests/golden/reference_nodes.npz) ----

def _ref_golden():
    return np.load(os.path.join(GOLD, "reference_nodes.npz"))


def _outside_sector_16(xyzi):
    az = np.degrees(np.arctan2(xyzi[:, 1].astype(np.float64), xyzi[:, 0].astype(np.float64)))
    return np.ascontiguousarray(xyzi[~((az > -8.5) & (az < 0.5))])


@pytest.mark.parametrize("seed", [0, 9])
def test_ground_node_matches_reference_golden(seed):
    z = _ref_golden()
    frame = _outside_sector_16(scans.generate(scans.config(2), 1, base_seed=seed)[0])
    assert hashlib.sha256(frame.tobytes()).hexdigest() == str(z[f"ground_seed{seed}_input_sha256"])
    exp, kept, _, _ = O.ground_node(O.view_of_xyzi(frame), GroundParams())
    e = np.stack([exp[n] for n in ("x", "y", "z", "pad", "intensity", "c1", "c2", "c3")], 1)
    assert kept == int(z[f"ground_seed{seed}_kept"])
    assert hashlib.sha256(np.ascontiguousarray(e).tobytes()).hexdigest() == str(z[f"ground_seed{seed}_sha256"])


@pytest.mark.parametrize("preset", ["our", "fsai", "simulation"])
def test_crop_matches_reference_lambda_golden(preset):
    """One verdict per point from the reference's compiled crop lambda, on clouds that sit on and one ulp around
    every threshold (including the x < 0, |y/x| huge points where libm's atan2f is one float off)."""
    z = _ref_golden()
    cloud, keep = z[f"crop_{preset}_cloud"], z[f"crop_{preset}_keep"].astype(bool)
    got = O.crop_mask(O.points32(cloud), PRESETS[preset]).astype(bool)
    assert np.array_equal(got, keep)
    assert 100 < keep.sum() < len(keep) - 100


@pytest.mark.parametrize("buffer", [True, False])
def test_detect_sequence_matches_reference_golden(buffer):
    from tests.util import TrackerReference
    z = _ref_golden()
    cfg = scans.config(1)
    d = cfg.detect
    ref = TrackerReference(False, buffer, d.cones_matching_dist_theshold, d.cone_position_extension_length)
    for fi, f in enumerate(scans.generate(cfg, 5, base_seed=40)):
        cl, _, _ = O.detect(O.view_of_xyzi(f), d, None, O.PCL_FAITHFUL)
        exp = np.array([(p[0], p[1]) for p in ref.update([(c["x"], c["y"]) for c in cl])[0]], np.float32).reshape(-1, 2)
        assert np.array_equal(exp.view(np.uint32), z[f"detect_buffer{int(buffer)}_frame{fi}"].view(np.uint32)), fi
