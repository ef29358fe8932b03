"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle, bit for bit."""
import numpy as np
import pytest

from cones_perception_b200 import api, scans
from cones_perception_b200.params import PRESETS, GroundParams
from cones_perception_b200.pointcloud2 import PointCloud2
from oracle import oracle as O
from tests.util import assert_frame_parity, boundary_cloud, oracle_stages, run_batch_with_taps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["fast", "general"])
def gpu(request):
    """Both back-half variants: the per-frame shared-memory kernel (which falls back to the
    general path on frames it cannot hold) and the general global-memory path forced on."""
    g = api.ConesGpu(max_points=1 << 22, max_frames=64, taps=True, back_mode=0 if request.param == "fast" else 2)
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
