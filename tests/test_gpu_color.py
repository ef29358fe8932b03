"""-m gpu: colour-path inputs through the C ABI.

* the range image against the reference's OWN to_image (tests/golden/cone_images.npz, produced by
  scripts/color_classifier_server.py:130-156 executed in the build container) — pinned parity;
* the box gather against the oracle restatement of get_reconstructed_cone
  (src/cone_detection.cpp:222-238), bit for bit and in cloud order."""
import os

import numpy as np
import pytest

from cones_perception_b200 import api, scans
from cones_perception_b200.pointcloud2 import PointCloud2, PointField
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CONE_WIDTH = 0.228


@pytest.fixture(scope="module")
def gpu():
    g = api.ConesGpu(max_points=1 << 20, max_frames=8, max_point_step=32)
    yield g
    g.close()


def _golden():
    z = np.load(os.path.join(GOLD, "cone_images.npz"))
    crops = np.load(os.path.join(GOLD, "cone_crops.npz"))
    return z, crops["points"], crops["lengths"], z["synth_points"], z["synth_lengths"]


def _offsets(lengths):
    return np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint32)


def test_raster_real_crops_match_reference(gpu):
    z, pts, lens, _, _ = _golden()
    img, flags = gpu.rasterize_crops(pts, _offsets(lens))
    assert not (flags & ~np.uint32(api.CONE_AMBIGUOUS)).any()
    assert np.array_equal(img, z["real_images"])           # 577 hand-labelled VLP-16 cone crops, byte for byte
    assert int((flags & api.CONE_AMBIGUOUS != 0).sum()) == 0


def test_raster_synthetic_crops_match_reference(gpu):
    z, _, _, pts, lens = _golden()
    img, flags = gpu.rasterize_crops(pts, _offsets(lens))
    raised = z["synth_raised"].astype(np.uint32)
    ok = raised == 0
    assert np.array_equal(img[ok], z["synth_images"][ok])
    assert not (flags[ok] & ~np.uint32(api.CONE_AMBIGUOUS)).any()
    # inputs on which numpy / interp1d raise: same failure class reported, image cleared
    assert np.all(flags[~ok] & raised[~ok]) and not img[~ok].any()
    assert int((~ok).sum()) >= 16


def test_raster_matches_oracle_on_random_crops(gpu):
    rng = np.random.default_rng(11)
    crops = []
    for i in range(300):
        r, az = rng.uniform(1.0, 30.0), rng.uniform(-np.pi, np.pi)
        n = int(rng.integers(1, 400))
        x = r * np.cos(az) + rng.uniform(-0.152, 0.152, n)
        y = r * np.sin(az) + rng.uniform(-0.152, 0.152, n)
        zz = rng.uniform(-0.26, 0.26, n) * min(r, 3.0) / 3.0
        crops.append(np.stack([x, y, zz, rng.uniform(0, 255, n)], 1).astype(np.float32))
    crops.append(np.zeros((0, 4), np.float32))                                  # empty: the service skips it
    crops.append(np.array([[np.nan, 1, 0, 5]], np.float32))                     # non-finite coordinate
    crops.append(np.array([[2, 1, 0, np.nan]], np.float32))                     # NaN intensity
    crops.append(np.array([[2, 1, 0, -1e-3]], np.float32))                      # below interp1d's range
    crops.append(np.array([[0, 0, 0, 5], [0, 0, 0, 9]], np.float32))            # the ground node's zero points
    img, flags = gpu.rasterize_crops(np.concatenate(crops), _offsets([len(c) for c in crops]))
    for i, c in enumerate(crops):
        eimg, efl = O.to_image(c)
        assert int(flags[i]) & ~api.CONE_AMBIGUOUS == efl, (i, flags[i], efl)
        if not flags[i] & api.CONE_AMBIGUOUS:
            assert np.array_equal(img[i], eimg), i
    assert int((flags & api.CONE_AMBIGUOUS != 0).sum()) <= 2


def _oracle_crops(frame_pts, centers, cone_width=CONE_WIDTH):
    out = [O.reconstruct_cone(frame_pts, float(c[0]), float(c[1]), cone_width) for c in centers]
    off = _offsets([len(c) for c in out])
    if off[-1] == 0:
        return off, np.zeros((0, 4), np.float32)
    flat = np.concatenate([np.stack([c["x"], c["y"], c["z"], c["intensity"]], 1) for c in out if len(c)])
    return off, flat.astype(np.float32)


def _extended_centers(frame, cfg):
    cl, _, _ = O.detect(O.view_of_xyzi(frame), cfg.detect, cfg.ground, O.CANONICAL)
    return np.array([O.extend(float(c["x"]), float(c["y"]), 0.05) for c in cl], np.float32)


def test_cone_crops_match_oracle_after_detect(gpu):
    cfg = scans.config(2)
    frame = scans.generate(cfg, 1, base_seed=4)[0]
    msg = PointCloud2.from_xyzi(frame)
    cl, _ = gpu.detect(msg, cfg.detect, cfg.ground)
    centers = _extended_centers(frame, cfg)
    assert len(centers) == len(cl) > 10
    eoff, epts = _oracle_crops(O.from_msg(O.view_of_xyzi(frame)), centers)
    # cloud passed again, and the cloud the detection call left on the device
    for kw in ({"msg": msg}, {"msg": None, "frame": 0}):
        off, pts = gpu.cone_crops(centers, CONE_WIDTH, **kw)
        assert np.array_equal(off, eoff)
        assert np.array_equal(pts.view(np.uint32), epts.view(np.uint32))
    assert eoff[-1] > 200
    img, counts, flags = gpu.cone_images(centers, CONE_WIDTH)
    assert np.array_equal(counts, np.diff(eoff))
    for i in range(len(centers)):
        eimg, efl = O.to_image(epts[eoff[i]:eoff[i + 1]])
        assert int(flags[i]) & ~api.CONE_AMBIGUOUS == efl
        if not flags[i] & api.CONE_AMBIGUOUS:
            assert np.array_equal(img[i], eimg), i


def test_cone_crops_many_overlapping_centres_and_batch_frames(gpu):
    cfg = scans.config(2)
    frames = scans.generate(cfg, 3, base_seed=20)
    gpu.set_host_input([PointCloud2.from_xyzi(f) for f in frames])
    rng = np.random.default_rng(3)
    for f in (2, 0):
        # 150 centres (three 64-centre chunks) dropped on random returns: boxes overlap, some hold ground rings
        pick = frames[f][rng.integers(0, len(frames[f]), 150)]
        centers = (pick[:, :2] + rng.normal(0, 0.05, (150, 2))).astype(np.float32)
        centers[7] = (np.nan, 0.0)
        centers[8] = (1e9, 1e9)
        eoff, epts = _oracle_crops(O.from_msg(O.view_of_xyzi(frames[f])), centers)
        off, pts = gpu.cone_crops(centers, CONE_WIDTH, frame=f, cap_points=int(eoff[-1]))
        assert np.array_equal(off, eoff) and off[8] == off[7] and off[9] == off[8]
        assert np.array_equal(pts.view(np.uint32), epts.view(np.uint32))
        with pytest.raises(api.ConesGpuError) as e:
            gpu.cone_crops(centers, CONE_WIDTH, frame=f, cap_points=int(eoff[-1]) - 1)
        assert e.value.status == api.CP_E_CAPACITY
    with pytest.raises(api.ConesGpuError) as e:
        gpu.cone_crops(centers, CONE_WIDTH, frame=3)
    assert e.value.status == api.CP_E_PARAM


def test_cone_crops_box_edges_exact(gpu):
    """Points one ulp inside / outside each face of the box, for centres whose double sums do and do not
    round when narrowed to float."""
    rng = np.random.default_rng(5)
    hw = float(np.float32(CONE_WIDTH)) / 1.5
    centers = rng.uniform(-20, 20, (40, 2)).astype(np.float32)
    rows = []
    for cx, cy in centers:
        for c, other, axis in ((cx, cy, 0), (cy, cx, 1)):
            for bound in (float(c) + hw, float(c) - hw):
                f = np.float32(bound)
                for v in (np.nextafter(f, np.float32(-np.inf)), f, np.nextafter(f, np.float32(np.inf))):
                    p = [v, other] if axis == 0 else [other, v]
                    rows.append(p + [rng.uniform(-0.3, 0.3), rng.uniform(0, 100)])
    cloud = np.array(rows, np.float32)
    cloud = cloud[rng.permutation(len(cloud))]
    eoff, epts = _oracle_crops(O.points32(cloud), centers)
    off, pts = gpu.cone_crops(centers, CONE_WIDTH, msg=PointCloud2.from_xyzi(cloud))
    assert np.array_equal(off, eoff) and np.array_equal(pts.view(np.uint32), epts.view(np.uint32))
    assert 0 < eoff[-1] < len(cloud)


@pytest.mark.parametrize("layout", ["pcl32", "odd"])
def test_cone_crops_other_layouts(gpu, layout):
    cfg = scans.config(1)
    frame = scans.generate(cfg, 1, base_seed=2)[0]
    centers = _extended_centers(frame, cfg)
    n = len(frame)
    if layout == "pcl32":      # what the ground_removal node publishes: x@0 y@4 z@8 intensity@16, step 32
        step, offs = 32, (0, 4, 8, 16)
    else:                      # unaligned fields: byte-wise loads
        step, offs = 23, (1, 5, 9, 14)
    raw = np.zeros((n, step), np.uint8)
    for k, o in enumerate(offs):
        raw[:, o:o + 4] = frame[:, k].copy().view(np.uint8).reshape(n, 4)
    msg = PointCloud2(data=raw.reshape(-1), width=n, height=1, point_step=step,
                      fields=[PointField(nm, o) for nm, o in zip(("x", "y", "z", "intensity"), offs)])
    eoff, epts = _oracle_crops(O.from_msg(O.view_of_xyzi(frame)), centers)
    off, pts = gpu.cone_crops(centers, CONE_WIDTH, msg=msg)
    assert np.array_equal(off, eoff) and np.array_equal(pts.view(np.uint32), epts.view(np.uint32))
    assert eoff[-1] > 50


def test_colour_path_degenerate_inputs(gpu):
    cloud = scans.generate(scans.config(1), 1, base_seed=0)[0]
    msg = PointCloud2.from_xyzi(cloud)
    off, pts = gpu.cone_crops(np.zeros((0, 2), np.float32), CONE_WIDTH, msg=msg)
    assert list(off) == [0] and len(pts) == 0
    img, counts, flags = gpu.cone_images(np.array([[500.0, 500.0]], np.float32), CONE_WIDTH, msg=msg)
    assert counts[0] == 0 and flags[0] == api.CONE_EMPTY and not img.any()
    empty = PointCloud2.from_xyzi(np.zeros((0, 4), np.float32))
    off, pts = gpu.cone_crops(np.array([[1.0, 1.0]], np.float32), CONE_WIDTH, msg=empty)
    assert list(off) == [0, 0]
    fresh = api.ConesGpu(max_points=1024)
    with pytest.raises(api.ConesGpuError) as e:
        fresh.cone_crops(np.array([[1.0, 1.0]], np.float32), CONE_WIDTH)
    assert e.value.status == api.CP_E_STATE
    fresh.close()
