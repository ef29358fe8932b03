"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle, bit for bit."""
import os

import numpy as np
import pytest

from cones_perception_b200 import api, scans
from cones_perception_b200.params import PRESETS, GroundParams
from cones_perception_b200.pointcloud2 import PointCloud2
from oracle import oracle as O
from tests.util import assert_frame_parity, boundary_cloud, oracle_stages, run_batch_with_taps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["fast", "general", "cluster"])
def gpu(request):
    """Pipeline variants: the per-frame shared-memory back half (falls back to the general path on
    frames it cannot hold), the general global-memory back half forced on, and the single-pass
    16-CTA-cluster front end in front of the fast back half."""
    g = api.ConesGpu(max_points=1 << 22, max_frames=64, taps=True, back_mode=3 if request.param == "general" else 0,
                     cluster_front=(request.param == "cluster"))
    yield g
    g.close()


@pytest.mark.parametrize("n,bits", [(0, 8), (1, 1), (31, 5), (2048, 8), (2049, 9), (100_000, 27), (1 << 20, 40),
                                    (300_000, 64)])
def test_radix_sort_matches_stable_argsort(gpu, n, bits):
    rng = np.random.default_rng(n + bits)
    keys = rng.integers(0, 1 << min(bits, 63), n, dtype=np.uint64) if n else np.zeros(0, np.uint64)
    if bits == 64 and n:
        keys |= rng.integers(0, 2, n, dtype=np.uint64) << np.uint64(63)
    # many duplicates in a second variant to exercise stability
    if n > 100:
        keys[::3] = keys[0]
    vals = np.arange(n, dtype=np.uint32)
    k, v = gpu.debug_sort(keys, vals, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, vals[order])


@pytest.mark.parametrize("cfg_idx", [1, 2])
def test_single_frame_configs_bit_exact(gpu, cfg_idx):
    cfg = scans.config(cfg_idx)
    frame = scans.generate(cfg, 1, base_seed=0)[0]
    ora = oracle_stages(frame, cfg.detect, cfg.ground)
    ctr, k_off, clusters, offs, taps = run_batch_with_taps(gpu, [frame], cfg.detect, cfg.ground)
    if cfg.ground is not None:
        assert np.array_equal(taps["low"][0].view(np.uint32), ora["low"].view(np.uint32)), "sector minima differ"
    assert_frame_parity(gpu, 0, ora, offs, taps, ctr, k_off, clusters)
    assert len(clusters) > 0


def test_detect_single_call_matches_oracle(gpu):
    cfg = scans.config(2)
    frame = scans.generate(cfg, 1, base_seed=3)[0]
    cl, ctr = gpu.detect(PointCloud2.from_xyzi(frame), cfg.detect, cfg.ground)
    exp, octr, _ = O.detect(O.view_of_xyzi(frame), cfg.detect, cfg.ground, O.CANONICAL)
    assert len(cl) == len(exp) and len(cl) > 0
    assert np.array_equal(cl.view(np.uint32), exp.view(np.uint32))
    assert ctr["n_cropped"] == octr.n_cropped and ctr["n_voxels"] == octr.n_voxels
    assert ctr["n_components"] == octr.n_components and ctr["key_bits"] == octr.key_bits


def test_ground_remove_node_output(gpu):
    cfg = scans.config(2)
    frame = scans.generate(cfg, 1, base_seed=5)[0]
    out, kept, low = gpu.ground_remove(PointCloud2.from_xyzi(frame), GroundParams())
    exp, ekept, elow, _ = O.ground_node(O.view_of_xyzi(frame), GroundParams())
    assert kept == ekept
    assert np.array_equal(low.view(np.uint32), elow.view(np.uint32))
    e = np.stack([exp[n] for n in ("x", "y", "z", "pad", "intensity", "c1", "c2", "c3")], 1)
    assert np.array_equal(out.view(np.uint32), e.view(np.uint32))


def test_uniform_and_ragged_batches(gpu):
    cfg = scans.config(3)
    frames = list(scans.generate(cfg, 6, base_seed=100))
    for ragged in (False, True):
        fr = frames
        if ragged:
            rng = np.random.default_rng(1)
            fr = [f[: int(n)] for f, n in zip(frames, rng.integers(1000, cfg.points_per_frame, len(frames)))]
            fr[2] = fr[2][:0]           # an empty frame in the middle
            fr[4] = fr[4][:2048]        # exactly one tile
        ctr, k_off, clusters, offs, taps = run_batch_with_taps(gpu, fr, cfg.detect, cfg.ground)
        for f, a in enumerate(fr):
            ora = oracle_stages(a, cfg.detect, cfg.ground)
            assert_frame_parity(gpu, f, ora, offs, taps, ctr, k_off, clusters)


def test_no_ground_removal_all_presets(gpu):
    cfg = scans.config(2)
    frame = scans.generate(cfg, 1, base_seed=9)[0]
    for name, d in PRESETS.items():
        ora = oracle_stages(frame, d, None)
        ctr, k_off, clusters, offs, taps = run_batch_with_taps(gpu, [frame], d, None)
        assert_frame_parity(gpu, 0, ora, offs, taps, ctr, k_off, clusters)


def test_adversarial_config5(gpu):
    cfg = scans.config(5)
    frame = scans.generate_config5(1, base_seed=0)[0]
    ora = oracle_stages(frame, cfg.detect, cfg.ground)
    ctr, k_off, clusters, offs, taps = run_batch_with_taps(gpu, [frame], cfg.detect, cfg.ground)
    assert_frame_parity(gpu, 0, ora, offs, taps, ctr, k_off, clusters)
    sizes = np.bincount(ora["labels"])
    assert sizes.max() > 5000, "the serpentine chain should form one deep component"
    assert (clusters["size"] == cfg.detect.max_cluster_size).any(), "the exactly-max component must be kept"
    assert not (clusters["size"] > cfg.detect.max_cluster_size).any()


@pytest.mark.parametrize("preset", ["our", "fsai", "simulation"])
@pytest.mark.parametrize("ground", [False, True])
def test_threshold_boundary_points_bit_exact(gpu, preset, ground):
    """Every guard-band fallback: points on / within ulps of each threshold must get the oracle's verdict."""
    d = PRESETS[preset]
    g = GroundParams() if ground else None
    frame = boundary_cloud(d, seed={"our": 1, "fsai": 2, "simulation": 3}[preset])
    ora = oracle_stages(frame, d, g)
    ctr, k_off, clusters, offs, taps = run_batch_with_taps(gpu, [frame], d, g)
    if ground:
        assert np.array_equal(taps["low"][0].view(np.uint32), ora["low"].view(np.uint32)), "sector minima differ"
    assert_frame_parity(gpu, 0, ora, offs, taps, ctr, k_off, clusters)


def test_golden_fixtures_on_gpu(gpu):
    import hashlib
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden")
    for idx, seed in ((1, 0), (2, 0), (2, 7), (4, 0), (5, 0)):
        z = np.load(os.path.join(gold, f"cfg{idx}_seed{seed}_golden.npz"))
        cfg = scans.config(idx)
        frame = scans.generate_config5(1, seed)[0] if idx == 5 else scans.generate(cfg, 1, seed)[0]
        assert hashlib.sha256(frame.tobytes()).hexdigest() == str(z["input_sha256"])
        with api.ConesGpu(max_points=len(frame), max_frames=1) as h:
            cl, ctr = h.detect(PointCloud2.from_xyzi(frame), cfg.detect, cfg.ground, cap=1 << 16)
        assert np.array_equal(cl.view(np.uint32), z["clusters"].view(np.uint32)), f"cfg{idx} seed {seed}"
        assert [int(ctr["n_cropped"]), int(ctr["n_voxels"]), int(ctr["n_components"]), int(ctr["n_clusters"]),
                int(ctr["key_bits"])] == z["counters"].tolist()[2:]


def _msg_with_layout(xyzi, point_step, offs, intensity=True, height=1, row_pad=0):
    """PointCloud2 with arbitrary point_step / field offsets / row padding around the same points."""
    from cones_perception_b200.pointcloud2 import PointField
    n = len(xyzi)
    width = n // height
    assert width * height == n
    row_step = width * point_step + row_pad
    raw = np.random.default_rng(5).integers(0, 255, height * row_step, dtype=np.uint8)  # junk between fields
    src = np.ascontiguousarray(xyzi, np.float32).view(np.uint8).reshape(n, 16)
    for r in range(height):
        for k, o in enumerate(offs):
            if k == 3 and not intensity:
                continue
            cols = (np.arange(width)[:, None] * point_step + o + np.arange(4)[None, :]) + r * row_step
            raw[cols.reshape(-1)] = src[r * width:(r + 1) * width, 4 * k:4 * k + 4].reshape(-1)
    names = ["x", "y", "z", "intensity"]
    fields = [PointField(names[k], offs[k]) for k in range(4 if intensity else 3)]
    return PointCloud2(data=raw, width=width, height=height, point_step=point_step, row_step=row_step, fields=fields)


@pytest.mark.parametrize("point_step,offs,intensity,height,row_pad", [
    (32, (0, 4, 8, 16), True, 1, 0),      # PCL PointXYZI layout: what the ground_removal node publishes
    (48, (4, 12, 20, 32), True, 1, 0),    # 4-byte aligned, reordered, padded (Ouster-like)
    (22, (0, 4, 8, 12), True, 1, 0),      # Velodyne-like unaligned point_step
    (19, (1, 6, 11, 15), True, 1, 0),     # everything unaligned
    (16, (0, 4, 8, 12), False, 1, 0),     # no intensity field: faked at offset 0 (aliases x)
    (32, (0, 4, 8, 16), True, 16, 64),    # organised cloud with row padding
])
def test_pointcloud2_layouts(gpu, point_step, offs, intensity, height, row_pad):
    cfg = scans.config(1)
    frame = scans.generate(cfg, 1, base_seed=77)[0][:29_984]   # divisible by 16 rows
    msg = _msg_with_layout(frame, point_step, offs, intensity, height, row_pad)
    for g in (None, GroundParams()):
        cl, ctr = gpu.detect(msg, cfg.detect, g)
        exp, octr, _ = O.detect(O.view_of_msg(msg, fake_missing_intensity=True), cfg.detect, g, O.CANONICAL)
        assert len(cl) == len(exp) and len(cl) > 0
        assert np.array_equal(cl.view(np.uint32), exp.view(np.uint32))
        assert int(ctr["n_voxels"]) == octr.n_voxels
    out, kept, low = gpu.ground_remove(msg, GroundParams())
    eout, ekept, elow, _ = O.ground_node(O.view_of_msg(msg, fake_missing_intensity=False), GroundParams())
    e = np.stack([eout[n] for n in ("x", "y", "z", "pad", "intensity", "c1", "c2", "c3")], 1)
    assert kept == ekept and np.array_equal(out.view(np.uint32), e.view(np.uint32))


def test_chained_nodes_equal_fused(gpu):
    """ground_removal node -> groundless_cloud topic -> cone_detection node (the reference's launch
    wiring) gives the same cones as the fused call, and both equal the oracle."""
    from cones_perception_b200.pointcloud2 import PointField
    cfg = scans.config(2)
    frame = scans.generate(cfg, 1, base_seed=78)[0]
    msg = PointCloud2.from_xyzi(frame)
    out32, kept, _ = gpu.ground_remove(msg, GroundParams())
    groundless = PointCloud2(data=out32.view(np.uint8).reshape(-1), width=len(frame), point_step=32,
                             fields=[PointField("x", 0), PointField("y", 4), PointField("z", 8),
                                     PointField("intensity", 16)])
    chained, _ = gpu.detect(groundless, cfg.detect, None)
    fused, _ = gpu.detect(msg, cfg.detect, GroundParams())
    exp, _, _ = O.detect(O.view_of_xyzi(frame), cfg.detect, GroundParams(), O.CANONICAL)
    assert np.array_equal(chained.view(np.uint32), exp.view(np.uint32))
    assert np.array_equal(fused.view(np.uint32), exp.view(np.uint32))


def test_capacity_and_parameter_errors(gpu):
    cfg = scans.config(1)
    frame = scans.generate(cfg, 1, base_seed=3)[0]
    msg = PointCloud2.from_xyzi(frame)
    with api.ConesGpu(max_points=1000, max_frames=1) as small:
        with pytest.raises(api.ConesGpuError) as e:
            small.detect(msg, cfg.detect, None)
        assert e.value.status == api.CP_E_CAPACITY
    with pytest.raises(api.ConesGpuError) as e:
        gpu.detect(msg, cfg.detect, None, cap=2)          # output buffer too small: never truncated
    assert e.value.status == api.CP_E_CAPACITY
    bad = PointCloud2.from_xyzi(frame)
    bad.fields = bad.fields[1:]                           # no x field
    with pytest.raises(api.ConesGpuError) as e:
        gpu.detect(bad, cfg.detect, None)
    assert e.value.status == api.CP_E_BADFIELD
    import dataclasses
    d = dataclasses.replace(cfg.detect, voxel_filter_leaf_size_x=0.0)
    with pytest.raises(api.ConesGpuError) as e:
        gpu.detect(msg, d, None)
    assert e.value.status == api.CP_E_PARAM
    # survivors overflow is reported, not truncated
    with api.ConesGpu(max_points=len(frame), max_frames=1, max_survivors=16, back_mode=3) as tiny:
        with pytest.raises(api.ConesGpuError) as e:
            tiny.detect(msg, cfg.detect, None)
        assert e.value.status == api.CP_E_CAPACITY


def test_empty_and_all_ground_frames(gpu):
    cfg = scans.config(2)
    empty = np.zeros((0, 4), np.float32)
    rng = np.random.default_rng(0)
    ground = np.zeros((5000, 4), np.float32)
    ground[:, 0] = rng.uniform(2, 6, 5000)
    ground[:, 1] = rng.uniform(-3, 3, 5000)
    ground[:, 2] = -0.6 + rng.normal(0, 0.002, 5000)
    real = scans.generate(cfg, 1, base_seed=9)[0]
    ctr, k_off, clusters, offs, taps = run_batch_with_taps(gpu, [empty, ground, real, empty], cfg.detect, cfg.ground)
    assert ctr["n_clusters"].tolist()[:2] == [0, 0] and ctr["n_clusters"][3] == 0 and ctr["n_clusters"][2] > 0
    assert ctr["n_cropped"][1] == 0
    for f, a in enumerate([empty, ground, real, empty]):
        assert_frame_parity(gpu, f, oracle_stages(a, cfg.detect, cfg.ground), offs, taps, ctr, k_off, clusters)


def test_zero_padding_survives_when_dmin_is_zero(gpu):
    """distance_treshold_min <= 0: the ground node's zero filler points survive the crop
    (src/ground_removal.cpp:79 + src/cone_detection.cpp:197) and join the voxel at the origin."""
    import dataclasses
    cfg = scans.config(2)
    frame = scans.generate(cfg, 1, base_seed=12)[0]
    d = dataclasses.replace(cfg.detect, distance_treshold_min=0.0, level_threshold=-5.0)
    ora = oracle_stages(frame, d, GroundParams())
    assert (ora["crop_index"] < 0).sum() > 1000                     # the filler points are really there
    ctr, k_off, clusters, offs, taps = run_batch_with_taps(gpu, [frame], d, GroundParams())
    assert_frame_parity(gpu, 0, ora, offs, taps, ctr, k_off, clusters)


def test_peer_gather_single_rank_roundtrip():
    """The peer-memory result path with world = 1: the publish kernel writes the packed cone list
    into the gather buffer and raises the sequence flag; the owner reads back what cp_batch_results
    returns.  (The 2-GPU path is exercised by `bench.py --gpus 2`, which asserts the same.)"""
    from cones_perception_b200.sharding import pack_words, unpack_gathered
    cfg = scans.config(3)
    frames = list(scans.generate(cfg, 4, base_seed=60))
    F, cap = len(frames), 4 * 64
    with api.ConesGpu(max_points=F * cfg.points_per_frame, max_frames=F) as h:
        words = pack_words(F, cap)
        handle = h.gather_create(1, words)
        assert len(handle) == 64
        for rep in range(3):                    # both parities of the double buffer
            h.set_host_input([PointCloud2.from_xyzi(f) for f in frames])
            h.run(cfg.detect, cfg.ground)
            ctr, off, cl = h.results()
            seq = h.gather_seq()
            assert seq == rep + 1
            h.gather_wait(seq)
            got = unpack_gathered(h.gather_read(seq, 1, words), F)
            assert [len(g) for g in got] == np.diff(off).tolist()
            assert np.array_equal(np.concatenate(got).view(np.uint32), cl.view(np.uint32))


def _random_scene(rng, n):
    """Unstructured random cloud: ground sheet with holes, blobs, thin lines, outliers, duplicates."""
    parts = []
    m = int(n * rng.uniform(0.3, 0.7))
    g = np.stack([rng.uniform(-12, 12, m), rng.uniform(-12, 12, m),
                  rng.uniform(-1.0, -0.3) + rng.normal(0, rng.uniform(0.001, 0.05), m)], 1)
    parts.append(g)
    for _ in range(int(rng.integers(3, 40))):
        c = rng.uniform(-9, 9, 3) * np.array([1, 1, 0.1])
        k = int(rng.integers(1, 400))
        parts.append(c + rng.normal(0, rng.uniform(0.01, 0.3), (k, 3)))
    for _ in range(int(rng.integers(0, 6))):
        a, b = rng.uniform(-8, 8, 3) * np.array([1, 1, 0.2]), rng.uniform(-8, 8, 3) * np.array([1, 1, 0.2])
        t = rng.uniform(0, 1, int(rng.integers(10, 300)))[:, None]
        parts.append(a + t * (b - a))
    parts.append(rng.uniform(-60, 60, (int(rng.integers(0, 50)), 3)))
    a = np.concatenate(parts)
    if len(a) > 10:
        a = np.concatenate([a, a[rng.integers(0, len(a), 10)]])      # exact duplicate points
    rng.shuffle(a)
    a = a[:n]
    out = np.zeros((len(a), 4), np.float32)
    out[:, :3] = a
    out[:, 3] = rng.uniform(0, 255, len(a))
    return out


@pytest.mark.parametrize("seed", range(int(os.environ.get("CONES_STRESS_SEEDS", "6"))))
def test_randomised_scenes_and_parameters(gpu, seed):
    """Random clouds x random (valid) parameter sets, several frames per batch, with and without
    ground removal: every stage must equal the oracle bit for bit."""
    import dataclasses
    rng = np.random.default_rng(1000 + seed)
    base = PRESETS[["our", "fsai", "simulation"][seed % 3]]
    d = dataclasses.replace(
        base,
        distance_treshold_max=float(rng.uniform(3, 14)), distance_treshold_min=float(rng.uniform(0.0, 2.0)),
        level_threshold=float(rng.uniform(-2.0, 0.2)), angle_threshold=float(rng.uniform(20, 200)),
        voxel_filter_leaf_size_x=float(rng.choice([0.03, 0.04, 0.05, 0.1])),
        voxel_filter_leaf_size_y=float(rng.choice([0.03, 0.04, 0.07])),
        voxel_filter_leaf_size_z=float(rng.choice([0.04, 0.08])),
        min_cluster_size=int(rng.integers(1, 5)), max_cluster_size=int(rng.integers(5, 600)))
    g = GroundParams(16, float(rng.choice([-0.1, -0.4, 0.0]))) if seed % 2 == 0 else None
    frames = [_random_scene(rng, int(rng.integers(500, 60_000))) for _ in range(4)]
    ctr, k_off, clusters, offs, taps = run_batch_with_taps(gpu, frames, d, g)
    for f, a in enumerate(frames):
        ora = oracle_stages(a, d, g)
        if g is not None:
            assert np.array_equal(taps["low"][f].view(np.uint32), ora["low"].view(np.uint32))
        assert_frame_parity(gpu, f, ora, offs, taps, ctr, k_off, clusters)


def test_row_skipping_is_exact_and_effective():
    """Pass 2 skips rows whose highest z is below every ground threshold.  The result must not depend
    on the skip (A/B against CONESGPU_ROWSKIP=0) and on a 64-beam scan most rows must be skipped."""
    cfg = scans.config(3)
    frames = list(scans.generate(cfg, 8, base_seed=70))
    msgs = [PointCloud2.from_xyzi(f) for f in frames]
    out = {}
    for skip in ("1", "0"):
        with api.ConesGpu(max_points=8 * cfg.points_per_frame, max_frames=8, env={"CONESGPU_ROWSKIP": skip}) as h:
            ctr, off, cl = h.detect_batch(msgs, cfg.detect, cfg.ground)
            out[skip] = (ctr.copy(), off.copy(), cl.copy(), h.last_rows_loaded())
    rows_total = 8 * cfg.points_per_frame // 32
    assert out["0"][3] == rows_total
    assert 0 < out["1"][3] < 0.3 * rows_total
    for a, b in zip(out["1"][:3], out["0"][:3]):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    for f, fr in enumerate(frames):
        exp, _, _ = O.detect(O.view_of_xyzi(fr), cfg.detect, cfg.ground, O.CANONICAL)
        got = out["1"][2][out["1"][1][f]:out["1"][1][f + 1]]
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


@pytest.mark.parametrize("ground", [True, False])
def test_pass2_inside_the_frame_kernel_equals_the_streaming_kernel(ground):
    """Pass 2 (ground verdicts + crop) runs inside the per-frame kernel by default (CONESGPU_FUSED_MASK unset; "2"
    forces it for small batches without ground removal) or as its own streaming kernel ("0").  Same cones, same
    counters (n_ground_kept included), same rows read — on a uniform batch, a ragged one with an empty frame and
    a one-tile frame, and the threshold-boundary cloud — and all equal to the oracle."""
    cfg = scans.config(3)
    g = cfg.ground if ground else None
    frames = list(scans.generate(cfg, 6, base_seed=900))
    rng = np.random.default_rng(5)
    ragged = [f[: int(n)] for f, n in zip(frames, rng.integers(1000, cfg.points_per_frame, len(frames)))]
    ragged[1] = ragged[1][:0]
    ragged[3] = ragged[3][:2048]
    ragged[4] = ragged[4][:33]
    batches = {"uniform": frames, "ragged": ragged, "boundary": [boundary_cloud(cfg.detect, seed=4)]}
    for name, fr in batches.items():
        msgs = [PointCloud2.from_xyzi(f) for f in fr]
        out = {}
        for fm in ("2", "0"):
            with api.ConesGpu(max_points=sum(len(f) for f in fr) + 1, max_frames=len(fr),
                              env={"CONESGPU_FUSED_MASK": fm}) as h:
                for _ in range(3):            # direct, graph capture, graph replay
                    ctr, off, cl = h.detect_batch(msgs, cfg.detect, g)
                out[fm] = (ctr.copy(), off.copy(), cl.copy(), h.last_rows_loaded(), h.last_launch_count())
        for a, b in zip(out["2"][:3], out["0"][:3]):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), name
        if name == "uniform":     # (a partly filled last tile: the streaming kernel also counts its padding rows)
            assert out["2"][3] == out["0"][3], (name, "rows read differ")
            if ground:
                assert out["2"][4] < out["0"][4], (name, "the fused path must need fewer launches")
            else:   # with the ground left in, these frames outgrow shared memory: both end on the general back half
                assert out["2"][4] <= out["0"][4], (name, "launch count")
        for f, a in enumerate(fr):
            exp, octr, _ = O.detect(O.view_of_xyzi(a), cfg.detect, g, O.CANONICAL)
            got = out["2"][2][out["2"][1][f]:out["2"][1][f + 1]]
            assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (name, f)
            assert int(out["2"][0]["n_ground_kept"][f]) == octr.n_ground_kept, (name, f)
            assert int(out["2"][0]["n_cropped"][f]) == octr.n_cropped, (name, f)


def test_frames_larger_than_one_mask_chunk():
    """The per-frame kernel walks a frame in chunks of 4 x CMAX rows (4096 rows = 131 072 points with the smallest
    budget): a 300 000-point frame crosses two chunk boundaries and must still come out like the oracle."""
    cfg = scans.config(3)
    parts = scans.generate(cfg, 3, base_seed=910)
    rng = np.random.default_rng(910)
    far = np.zeros((170_000, 4), np.float32)           # returns beyond distance_treshold_max, at and above the ground
    ang, rad = rng.uniform(-np.pi, np.pi, len(far)), rng.uniform(30, 60, len(far))
    far[:, 0], far[:, 1] = rad * np.cos(ang), rad * np.sin(ang)
    far[:, 2] = np.where(rng.random(len(far)) < 0.8, -0.6 + rng.normal(0, 0.004, len(far)), rng.uniform(-0.4, 2.0, len(far)))
    clouds = {"few survivors": np.ascontiguousarray(np.concatenate([parts[0][:70_001], far, parts[0][70_001:]])),
              "three scans in one (overflows the small budgets)": np.ascontiguousarray(np.concatenate(list(parts))[:300_000])}
    for name, big in clouds.items():
        exp, octr, _ = O.detect(O.view_of_xyzi(big), cfg.detect, cfg.ground, O.CANONICAL)
        for mode in (0, 2):
            with api.ConesGpu(max_points=len(big), max_frames=1, back_mode=mode) as h:
                cl, ctr = h.detect(PointCloud2.from_xyzi(big), cfg.detect, cfg.ground, cap=1 << 16)
            assert np.array_equal(cl.view(np.uint32), exp.view(np.uint32)), (name, mode)
            assert int(ctr["n_ground_kept"]) == octr.n_ground_kept and int(ctr["n_cropped"]) == octr.n_cropped


def test_gather_wired_after_a_graph_was_captured():
    """A handle that has already replayed a run from a CUDA graph gets its peer gather wired afterwards
    (warm-up, then setup): the graph must be dropped, or the publish kernel would never run again."""
    from cones_perception_b200.sharding import pack_words, unpack_gathered
    import torch
    cfg = scans.config(3)
    a = scans.generate(cfg, 3, base_seed=920)
    F, N = a.shape[0], a.shape[1]
    dev = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    with api.ConesGpu(max_points=F * N, max_frames=F) as h:
        h.set_device_input(dev.data_ptr(), np.full(F, N, np.uint32), keep=dev)
        for _ in range(3):                      # direct, capture, replay
            h.run(cfg.detect, cfg.ground)
            ctr, off, cl = h.results()
        words = pack_words(F, 4 * 64)
        h.gather_create(1, words)
        for rep in range(3):
            h.run(cfg.detect, cfg.ground)
            h.sync()
            seq = h.gather_seq()
            assert seq == rep + 1
            h.gather_wait(seq, timeout_ms=2000)
            got = unpack_gathered(h.gather_read(seq, 1, words), F)
            assert np.array_equal(np.concatenate(got).view(np.uint32), cl.view(np.uint32))


def test_gather_slot_overflow_is_reported():
    """More cones than the gather slot holds: never truncated silently — the publishing rank's cp_sync and the
    gathering rank's cp_gather_wait both fail with CP_E_CAPACITY."""
    from cones_perception_b200.sharding import pack_words
    cfg = scans.config(3)
    frames = list(scans.generate(cfg, 4, base_seed=60))
    F = len(frames)
    with api.ConesGpu(max_points=F * cfg.points_per_frame, max_frames=F) as h:
        with pytest.raises(api.ConesGpuError):
            h.gather_create(1, 4)               # cannot even hold the offsets
        words = pack_words(F, 8)                # room for 8 cones; the batch has ~100
        h.gather_create(1, words)
        h.set_host_input([PointCloud2.from_xyzi(f) for f in frames])
        h.run(cfg.detect, cfg.ground)
        with pytest.raises(api.ConesGpuError) as e:
            h.sync()
        assert e.value.status == api.CP_E_CAPACITY
        with pytest.raises(api.ConesGpuError) as e:
            h.gather_wait(h.gather_seq(), timeout_ms=2000)
        assert e.value.status == api.CP_E_CAPACITY


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_every_back_half_variant_gives_the_same_cones(mode):
    """All four back-half variants (three shared-memory budgets + the general path), forced one by one,
    on a batch whose frames fit every budget."""
    cfg = scans.config(3)
    frames = list(scans.generate(cfg, 5, base_seed=80))
    with api.ConesGpu(max_points=5 * cfg.points_per_frame, max_frames=5, back_mode=mode) as h:
        ctr, off, cl = h.detect_batch([PointCloud2.from_xyzi(f) for f in frames], cfg.detect, cfg.ground)
        single, _ = h.detect(PointCloud2.from_xyzi(frames[0]), cfg.detect, cfg.ground)   # few-frame (512-thread) path
    for f, fr in enumerate(frames):
        exp, octr, _ = O.detect(O.view_of_xyzi(fr), cfg.detect, cfg.ground, O.CANONICAL)
        got = cl[off[f]:off[f + 1]]
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (mode, f)
        assert int(ctr["n_cropped"][f]) == octr.n_cropped and int(ctr["n_voxels"][f]) == octr.n_voxels
        assert int(ctr["n_components"][f]) == octr.n_components and int(ctr["key_bits"][f]) == octr.key_bits
    assert np.array_equal(single.view(np.uint32), cl[off[0]:off[1]].view(np.uint32))


@pytest.mark.parametrize("mode", [0, 3])
def test_graph_replay_follows_the_data(mode):
    """A repeated identical run is captured into a CUDA graph on its second sighting and replayed afterwards.
    The replays must read the input afresh: the device buffer is overwritten in place between runs and every
    run must match the oracle of what is in the buffer at that time (per-frame kernel and general back half)."""
    import torch
    cfg = scans.config(3)
    a = scans.generate(cfg, 3, base_seed=300)
    b = scans.generate(cfg, 3, base_seed=400)
    F, N = a.shape[0], a.shape[1]
    dev = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    exp = {}
    for name, fr in (("a", a), ("b", b)):
        exp[name] = [O.detect(O.view_of_xyzi(f), cfg.detect, cfg.ground, O.CANONICAL)[0] for f in fr]
    with api.ConesGpu(max_points=F * N, max_frames=F, back_mode=mode) as h:
        h.set_device_input(dev.data_ptr(), np.full(F, N, np.uint32), keep=dev)
        for rep, name in enumerate(["a", "a", "a", "b", "b", "a", "b"]):   # direct, capture, replay, replays on new data
            dev.copy_(torch.from_numpy(np.ascontiguousarray(a if name == "a" else b)))
            torch.cuda.synchronize()
            h.run(cfg.detect, cfg.ground)
            ctr, off, cl = h.results()
            for f in range(F):
                assert np.array_equal(cl[off[f]:off[f + 1]].view(np.uint32), exp[name][f].view(np.uint32)), (rep, name, f)


def test_exact_path_atan2f_is_the_references(gpu):
    """The guard-band fallback of the kernels must be libm's atan2f as the reference calls it (fdlibm's routine,
    not correctly rounded), bit for bit: 1.2 M pairs against the oracle's restatement (itself pinned to libm)."""
    from tests.test_oracle import atan2_cases
    y, x = atan2_cases(seed=1)
    got = gpu.debug_atan2f(y, x)
    import ctypes as C
    f = O.lib().orc_atan2f
    exp = np.array([f(a, b) for a, b in zip(y.tolist(), x.tolist())], np.float32)
    same = (got.view(np.uint32) == exp.view(np.uint32)) | (np.isnan(got) & np.isnan(exp))
    assert same.all(), (y[~same][:5], x[~same][:5], got[~same][:5], exp[~same][:5])


# ---- the CUDA path against goldens produced by the reference's own compiled node sources ----

def _ref_golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_nodes.npz"))


@pytest.mark.parametrize("seed", [0, 9])
def test_ground_remove_matches_reference_node_golden(gpu, seed):
    """cp_ground_remove against what the real ground_removal node published (sha256 of the 32-byte cloud)."""
    import hashlib
    from tests.test_oracle import _outside_sector_16
    z = _ref_golden()
    frame = _outside_sector_16(scans.generate(scans.config(2), 1, base_seed=seed)[0])
    assert hashlib.sha256(frame.tobytes()).hexdigest() == str(z[f"ground_seed{seed}_input_sha256"])
    out, kept, _ = gpu.ground_remove(PointCloud2.from_xyzi(frame), GroundParams())
    assert kept == int(z[f"ground_seed{seed}_kept"])
    assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == str(z[f"ground_seed{seed}_sha256"])


@pytest.mark.parametrize("preset", ["our", "fsai", "simulation"])
def test_crop_matches_reference_lambda_golden(gpu, preset):
    """The crop kernel's keep bits against the verdicts of the reference's compiled crop lambda on the
    threshold-boundary clouds (fast paths + guard-band fallbacks, libm atan2f semantics included)."""
    z = _ref_golden()
    cloud, keep = z[f"crop_{preset}_cloud"], z[f"crop_{preset}_keep"].astype(bool)
    ctr, k_off, clusters, offs, taps = run_batch_with_taps(gpu, [cloud], PRESETS[preset], None)
    got = np.zeros(len(cloud), bool)
    got[taps["crop_index"][offs["c_off"][0]:offs["c_off"][1]]] = True
    assert np.array_equal(got, keep)


@pytest.mark.parametrize("buffer", [True, False])
def test_detect_sequence_matches_reference_node_golden(gpu, buffer):
    """cp_detect + the tracker restatement against the clouds the real cone_detection node published.  The node's
    voxel means use PCL's (implementation-defined) summation order, the CUDA path the canonical one: same cones,
    centroids within north_star's 1e-5 m."""
    from tests.util import TrackerReference
    z = _ref_golden()
    cfg = scans.config(1)
    d = cfg.detect
    ref = TrackerReference(False, buffer, d.cones_matching_dist_theshold, d.cone_position_extension_length)
    for fi, f in enumerate(scans.generate(cfg, 5, base_seed=40)):
        cl, _ = gpu.detect(PointCloud2.from_xyzi(f), d, None)
        got = np.array([(p[0], p[1]) for p in ref.update([(c["x"], c["y"]) for c in cl])[0]], np.float32).reshape(-1, 2)
        exp = z[f"detect_buffer{int(buffer)}_frame{fi}"]
        assert got.shape == exp.shape, fi
        if len(exp):
            a, b = got[np.lexsort(got.T)], exp[np.lexsort(exp.T)]
            assert np.allclose(a, b, rtol=0, atol=1e-5), fi


@pytest.mark.parametrize("mode", [1, 2])
def test_two_devices_in_one_process(mode):
    """One process holding handles on two GPUs (kernel function attributes are per device): the larger
    shared-memory variants of the per-frame kernel must work on both."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    cfg = scans.config(3)
    frames = list(scans.generate(cfg, 3, base_seed=500))
    exp = [O.detect(O.view_of_xyzi(f), cfg.detect, cfg.ground, O.CANONICAL)[0] for f in frames]
    for dev in (0, 1, 0):
        with api.ConesGpu(max_points=3 * cfg.points_per_frame, max_frames=3, device=dev, back_mode=mode) as h:
            ctr, off, cl = h.detect_batch([PointCloud2.from_xyzi(f) for f in frames], cfg.detect, cfg.ground)
        for f in range(3):
            assert np.array_equal(cl[off[f]:off[f + 1]].view(np.uint32), exp[f].view(np.uint32)), (dev, f)


def test_interleaved_handles_on_two_devices():
    """Two handles on different GPUs used alternately from one thread: every entry point that may launch
    (cp_sync publishes results with a kernel) must select its handle's device first."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    cfg = scans.config(2)
    fa, fb = scans.generate(cfg, 1, base_seed=600)[0], scans.generate(cfg, 1, base_seed=601)[0]
    ea = O.detect(O.view_of_xyzi(fa), cfg.detect, cfg.ground, O.CANONICAL)[0]
    eb = O.detect(O.view_of_xyzi(fb), cfg.detect, cfg.ground, O.CANONICAL)[0]
    with api.ConesGpu(max_points=len(fa), device=0) as a, api.ConesGpu(max_points=len(fb), device=1) as b:
        for _ in range(4):
            a.set_host_input([PointCloud2.from_xyzi(fa)])
            b.set_host_input([PointCloud2.from_xyzi(fb)])
            a.run(cfg.detect, cfg.ground)
            b.run(cfg.detect, cfg.ground)
            a.sync()                         # current device is 1 here
            _, _, ca = a.results()
            b.sync()
            _, _, cb = b.results()
            assert np.array_equal(ca.view(np.uint32), ea.view(np.uint32))
            assert np.array_equal(cb.view(np.uint32), eb.view(np.uint32))


@pytest.mark.parametrize("threads,nt", [("1", "0"), ("1", "1"), ("2", "1"), ("4", "0"), ("4", "1"), ("7", "1")])
def test_pageable_staging_variants(threads, nt):
    """Pageable host clouds go through the pinned staging ring: serial or shared with helper threads, plain or
    non-temporal stores, one chunk (2 MB frame) or several (config 4's 42 MB frame exceeds the 32 MB ring half)."""
    env = {"CONESGPU_STAGE_THREADS": threads, "CONESGPU_STAGE_NT": nt}
    for idx in (2, 4):
        cfg = scans.config(idx)
        frame = scans.generate(cfg, 1, base_seed=11)[0]
        frame = np.ascontiguousarray(frame[3:])            # odd point count, source not 64-byte aligned
        exp, _, _ = O.detect(O.view_of_xyzi(frame), cfg.detect, cfg.ground, O.CANONICAL)
        with api.ConesGpu(max_points=len(frame), max_frames=1, env=env) as h:
            for _ in range(2):
                cl, _ = h.detect(PointCloud2.from_xyzi(frame), cfg.detect, cfg.ground, cap=1 << 16)
                assert np.array_equal(cl.view(np.uint32), exp.view(np.uint32)), (idx, threads, nt)


@pytest.mark.parametrize("mode", [0, 3])
def test_pairs_tested_counter(mode):
    """cp_last_pairs (SURVEY §5 "pairs tested"): both union kernels count the candidate pairs they visit and the
    distance tests they make.  Every link of the union-find needs one tested pair, so tested >= V - components;
    the early exit for voxels already under the same root keeps tested well below visited on a scan of cones."""
    cfg = scans.config(3)
    frames = list(scans.generate(cfg, 4, base_seed=21))
    with api.ConesGpu(max_points=4 * cfg.points_per_frame, max_frames=4, back_mode=mode) as h:
        ctr, off, cl = h.detect_batch([PointCloud2.from_xyzi(f) for f in frames], cfg.detect, cfg.ground)
        visited, tested = h.last_pairs()
        V, comps = int(ctr["n_voxels"].sum()), int(ctr["n_components"].sum())
        assert visited >= tested >= V - comps > 0
        assert visited <= V * V
        h.detect_batch([PointCloud2.from_xyzi(f) for f in frames], cfg.detect, cfg.ground)   # counters restart per run
        v2, t2 = h.last_pairs()
        assert v2 <= visited * 2 and t2 > 0


@pytest.mark.parametrize("case", ["cfg2_ground", "cfg1_no_ground", "pcl32_no_ground", "tiny", "odd_sizes", "overflow"])
def test_single_frame_in_one_launch_equals_the_multi_launch_path(case):
    """A single frame is ONE kernel launch (single_frame.cuh: 16-CTA cluster front end, back half and result publish
    fused): same cones and counters as the multi-launch path (CONESGPU_SINGLE=0) and as the oracle, on the direct
    run, the graph capture and the replay; frames that outgrow the shared-memory budget fall back and still agree."""
    cfg = scans.config(2)
    g = cfg.ground
    d = cfg.detect
    if case == "cfg2_ground":
        clouds = [scans.generate(cfg, 1, base_seed=11)[0]]
    elif case == "cfg1_no_ground":
        cfg = scans.config(1)
        d, g = cfg.detect, None
        clouds = [scans.generate(cfg, 1, base_seed=12)[0]]
    elif case == "pcl32_no_ground":     # what the detection node receives from the ground node: 32-byte PCL points
        cfg = scans.config(1)
        d, g = cfg.detect, None
        clouds = [scans.generate(cfg, 1, base_seed=13)[0]]
    elif case == "tiny":
        f = scans.generate(cfg, 1, base_seed=14)[0]
        clouds = [f[:1], f[:33], f[:2048], f[:4097]]
    elif case == "odd_sizes":
        f = scans.generate(cfg, 1, base_seed=15)[0]
        clouds = [f[:65535], f[:100001], np.concatenate([f, f])[:262144], np.concatenate([f, f])[:262145]]
    else:                               # ground left in: C and V outgrow every shared-memory budget
        g = None
        clouds = [scans.generate(scans.config(3), 1, base_seed=16)[0]]
    for cloud in clouds:
        if case == "pcl32_no_ground":
            msg = _msg_with_layout(cloud, 32, (0, 4, 8, 16))
        else:
            msg = PointCloud2.from_xyzi(cloud)
        out = {}
        for single in ("1", "0"):
            with api.ConesGpu(max_points=max(len(cloud), 1), max_frames=1, max_point_step=32,
                              env={"CONESGPU_SINGLE": single}) as h:
                runs = [h.detect(msg, d, g, cap=1 << 16) for _ in range(4)]
                for cl, ctr in runs[1:]:
                    assert np.array_equal(cl.view(np.uint32), runs[0][0].view(np.uint32))
                    assert ctr.tobytes() == runs[0][1].tobytes()
                out[single] = (runs[0][0], runs[0][1], h.last_launch_count())
        assert np.array_equal(out["1"][0].view(np.uint32), out["0"][0].view(np.uint32)), (case, len(cloud))
        assert out["1"][1].tobytes() == out["0"][1].tobytes(), (case, len(cloud), out["1"][1], out["0"][1])
        exp, octr, _ = O.detect(O.view_of_xyzi(cloud), d, g, O.CANONICAL)
        assert np.array_equal(out["1"][0].view(np.uint32), exp.view(np.uint32)), (case, len(cloud))
        assert int(out["1"][1]["n_ground_kept"]) == octr.n_ground_kept and int(out["1"][1]["n_cropped"]) == octr.n_cropped
        if case not in ("overflow",) and len(cloud) <= 262144:
            assert out["1"][2] == 1, (case, len(cloud), "a single frame must be one launch")
            assert out["0"][2] > 1


def test_results_do_not_depend_on_batching_or_frame_order():
    """Size-independent properties of the path (frames are independent units): a frame's cones are the same whether
    it runs alone (one-launch kernel), inside a batch, or inside the same batch in another order; running a batch
    twice changes nothing (idempotence); within every frame the list is sorted by size descending, then min index."""
    cfg = scans.config(3)
    F = 48
    frames = list(scans.generate(cfg, F, base_seed=300))
    rng = np.random.default_rng(9)
    perm = rng.permutation(F)
    msgs = [PointCloud2.from_xyzi(f) for f in frames]
    with api.ConesGpu(max_points=F * cfg.points_per_frame, max_frames=F) as h:
        ctr, off, cl = h.detect_batch(msgs, cfg.detect, cfg.ground)
        ctr2, off2, cl2 = h.detect_batch(msgs, cfg.detect, cfg.ground)
        assert np.array_equal(off, off2) and cl.tobytes() == cl2.tobytes() and ctr.tobytes() == ctr2.tobytes()
        pctr, poff, pcl = h.detect_batch([msgs[i] for i in perm], cfg.detect, cfg.ground)
        for k, f in enumerate(perm):
            assert pcl[poff[k]:poff[k + 1]].tobytes() == cl[off[f]:off[f + 1]].tobytes(), f
            assert pctr[k].tobytes() == ctr[f].tobytes()
        for f in range(F):
            c = cl[off[f]:off[f + 1]]
            key = list(zip((-c["size"].astype(np.int64)).tolist(), c["min_index"].tolist()))
            assert key == sorted(key), f
            assert (c["size"] >= cfg.detect.min_cluster_size).all() and (c["size"] <= cfg.detect.max_cluster_size).all()
    with api.ConesGpu(max_points=cfg.points_per_frame, max_frames=1) as h1:
        for f in (0, 17, F - 1):
            one, c1 = h1.detect(msgs[f], cfg.detect, cfg.ground)
            assert one.tobytes() == cl[off[f]:off[f + 1]].tobytes() and c1.tobytes() == ctr[f].tobytes()


@pytest.mark.parametrize("ground", [False, True])
def test_general_front_end_parks_sparse_tiles_and_resolves_dense_ones_on_the_spot(ground):
    """mask_compact_kernel (general back half) parks a tile with at most 1024 survivors in shared memory and
    resolves its look-back one tile later; a tile with more survivors resolves on the spot and stores from
    registers.  A batch that interleaves both kinds of tiles — whole 2048-point tiles that survive the crop next to
    tiles where nothing or little does, with and without the pad record of the ground node — must come out of the
    general path exactly like the oracle, frame by frame, also when the run is replayed from a graph."""
    cfg = scans.config(3)
    d = cfg.detect
    g = cfg.ground if ground else None
    rng = np.random.default_rng(77 + int(ground))
    frames = []
    for f in range(3):
        base = scans.generate(cfg, 1, base_seed=400 + f)[0]
        n_dense = 3 * 2048 + 700                       # three tiles and a bit in which every point survives
        dense = np.zeros((n_dense, 4), np.float32)
        ang = rng.uniform(-0.6, 0.6, n_dense)
        rad = rng.uniform(3.0, 9.0, n_dense)
        dense[:, 0], dense[:, 1] = rad * np.cos(ang), rad * np.sin(ang)
        dense[:, 2] = rng.uniform(0.3, 0.6, n_dense)    # above the ground band, below the level threshold's cut
        dense[:, 3] = rng.uniform(0, 100, n_dense)
        far = np.zeros((2 * 2048 + 100, 4), np.float32)   # beyond distance_treshold_max: tiles without a survivor
        fa, fr_ = rng.uniform(-np.pi, np.pi, len(far)), rng.uniform(30, 60, len(far))
        far[:, 0], far[:, 1], far[:, 2] = fr_ * np.cos(fa), fr_ * np.sin(fa), rng.uniform(-0.5, 2.0, len(far))
        for c in range(6):                              # six cones' worth of returns scattered through those tiles
            ca, cr = 1.2 + 0.25 * c, 4.0 + 0.7 * c
            idx = rng.choice(len(far), 40, replace=False)
            far[idx, 0] = cr * np.cos(ca) + rng.uniform(-0.08, 0.08, 40)
            far[idx, 1] = cr * np.sin(ca) + rng.uniform(-0.08, 0.08, 40)
            far[idx, 2] = rng.uniform(0.0, 0.3, 40)
        cut = 2048 * (5 + f) + 13 * f                   # not tile-aligned in frames 1, 2
        frames.append(np.ascontiguousarray(np.concatenate([far, base[:cut], dense, far[::-1], base[cut:cut + 40_000]])))
    clouds = [PointCloud2.from_xyzi(fr) for fr in frames]
    with api.ConesGpu(max_points=sum(len(fr) for fr in frames), max_frames=3, back_mode=3) as h:
        for rep in range(3):                            # direct run, graph capture, replay
            ctr, off, cl = h.detect_batch(clouds, d, g)
            for f, fr in enumerate(frames):
                exp, octr, _ = O.detect(O.view_of_xyzi(fr), d, g, O.CANONICAL)
                assert int(ctr["n_cropped"][f]) == octr.n_cropped, (f, rep)
                assert int(ctr["n_voxels"][f]) == octr.n_voxels, (f, rep)
                assert np.array_equal(cl[off[f]:off[f + 1]].view(np.uint32), exp.view(np.uint32)), (f, rep)
                if g is not None:
                    assert int(ctr["n_ground_kept"][f]) == octr.n_ground_kept, (f, rep)
            assert int(ctr["n_cropped"].max()) > 3 * 2048 and int(off[-1]) >= 6
