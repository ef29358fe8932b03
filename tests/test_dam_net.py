"""The colour classifier: models/dam_net/dam_net.tflite as evaluated by the reference's service
(scripts/color_classifier_server.py:66-71, :108-120).

PARITY UNPINNED for the network: TensorFlow-Lite cannot be installed here, so there is no interpreter output to
compare with.  What holds the implementation instead:
  * the flatbuffer is read by two independent readers (Python, cones_perception_b200/tflite_model.py; C++ inside
    libconesgpu, csrc/tflite_reader.hpp) that must extract the same graph;
  * the numpy forward pass (oracle/dam_net_ref.py, TFLite's reference-kernel arithmetic) is pinned STATISTICALLY on
    the 577 human-labelled crops the reference ships (cones_clouds/cones.pkl, `color` column): label agreement at
    the rates asserted below, confusion matrices in the assertion messages / DESIGN.md;
  * the CUDA kernel must reproduce that forward pass bit for bit on the logits.
"""
import os

import numpy as np
import pytest

from cones_perception_b200 import tflite_model as T
from oracle import dam_net_ref as D

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def model_bytes() -> bytes:
    z = np.load(os.path.join(GOLD, "dam_net_model.npz"))
    return z["tflite"].tobytes()


def real_images():
    return np.load(os.path.join(GOLD, "cone_images.npz"))["real_images"], \
        np.load(os.path.join(GOLD, "cone_crops.npz"))["colors"].astype(int)


def confusion(graph, imgs, labels, scale=1.0):
    thr, arg = np.zeros((4, 4), int), np.zeros((4, 4), int)
    for im, lab in zip(imgs, labels):
        x = np.clip(im.astype(np.float32) * scale, 0, 255).astype(np.uint8)
        p, _ = D.forward(graph, x)
        thr[lab, D.decide(p)] += 1
        arg[lab, int(np.argmax(p)) + 1] += 1
    return thr, arg


def test_reader_extracts_the_dam_net_graph():
    g = T.load(model_bytes())
    assert [o.kind for o in g.ops] == ["CONV_2D", "MAX_POOL_2D", "CONV_2D", "MAX_POOL_2D", "MUL", "ADD", "RESHAPE",
                                       "FULLY_CONNECTED", "SOFTMAX"]
    assert g.tensors[g.inputs[0]].shape == (1, 15, 12, 1) and g.tensors[g.outputs[0]].shape == (1, 3)
    conv1, conv2, fc = g.ops[0], g.ops[2], g.ops[7]
    assert g.tensors[conv1.inputs[1]].data.shape == (16, 3, 3, 1) and conv1.options["act"] == "RELU"
    assert g.tensors[conv2.inputs[1]].data.shape == (32, 3, 3, 16) and conv2.options["padding"] == "VALID"
    assert g.tensors[fc.inputs[1]].data.shape == (3, 64)
    for bad in (b"", b"\x00" * 64, model_bytes()[:100]):
        with pytest.raises((T.UnsupportedModel, Exception)):
            T.load(bad)


def test_tflite_tensors_are_the_keras_models_variables():
    """Second artifact of the reference: the Keras SavedModel beside the .tflite file.  The tensors the reader
    extracts must be its variables byte for byte (kernels transposed HWIO -> OHWI, dense (64,3) -> (3,64)), and
    the MUL / ADD constants must be its batch normalisation folded with epsilon 0.001 (keras_metadata.pb):
    scale = gamma / sqrt(var + eps), shift = beta - mean * scale.  This pins which tensor plays which role."""
    g = T.load(model_bytes())
    kv = np.load(os.path.join(GOLD, "dam_net_model.npz"))["keras_variables"].tobytes()
    f = np.frombuffer(kv, np.float32)
    pos = 0

    def take(*shape):
        nonlocal pos
        n = int(np.prod(shape))
        a = f[pos:pos + n].reshape(shape)
        pos += n
        return a

    k1, b1 = take(3, 3, 1, 16), take(16)
    k2, b2 = take(3, 3, 16, 32), take(32)
    gamma, beta, mean, var = take(32), take(32), take(32), take(32)
    kd, bd = take(64, 3), take(3)
    assert pos * 4 == len(kv)
    conv1, conv2, mul, add, fc = g.ops[0], g.ops[2], g.ops[4], g.ops[5], g.ops[7]
    tens = lambda i: g.tensors[i].data
    assert np.array_equal(tens(conv1.inputs[1]), np.transpose(k1, (3, 0, 1, 2))) and np.array_equal(tens(conv1.inputs[2]), b1)
    assert np.array_equal(tens(conv2.inputs[1]), np.transpose(k2, (3, 0, 1, 2))) and np.array_equal(tens(conv2.inputs[2]), b2)
    assert np.array_equal(tens(fc.inputs[1]), kd.T) and np.array_equal(tens(fc.inputs[2]), bd)
    scale = gamma / np.sqrt(var + np.float32(1e-3))
    shift = beta - mean * scale
    assert np.max(np.abs(tens(mul.inputs[1]) - scale)) <= 2e-7 and np.max(np.abs(tens(add.inputs[1]) - shift)) <= 2e-7
    assert (var > 0).all()


def test_forward_pass_regression_pin():
    g = T.load(model_bytes())
    z = np.load(os.path.join(GOLD, "dam_net_model.npz"))
    imgs, _ = real_images()
    for i in range(0, len(imgs), 7):
        p, l = D.forward(g, imgs[i])
        assert np.array_equal(l.view(np.uint32), z["real_logits"][i].view(np.uint32))
        assert np.array_equal(p.view(np.uint32), z["real_probs"][i].view(np.uint32))
        assert abs(float(p.sum()) - 1.0) < 1e-6


def test_label_agreement_on_the_577_recorded_crops():
    """Statistical pin.  With the images exactly as the reference's service builds them today (identity intensity
    mapping, color_classifier_server.py:36) the network's argmax agrees with the human label on >= 89 % of the
    crops and the thresholded decision (>= 0.8, :116-120) on 60 %: most blue and orange cones come out "unknown".
    With the intensity mapping the file keeps commented out (:35, [0,100] -> [0,255], evidently the one the network
    was trained with) the same forward pass agrees on >= 99 % (argmax) / >= 96 % (thresholded).  A wrong layer
    order, weight layout or activation cannot reach those rates."""
    g = T.load(model_bytes())
    imgs, labels = real_images()
    thr, arg = confusion(g, imgs, labels, 1.0)
    n = len(imgs)
    assert np.trace(arg) / n >= 0.89, f"argmax confusion (rows = label 0..3, cols = decision):\n{arg}"
    assert 0.58 <= np.trace(thr) / n <= 0.63, f"thresholded confusion:\n{thr}"
    assert thr[1, 1] == (labels == 1).sum(), "every yellow cone is recognised even with the identity mapping"
    thr2, arg2 = confusion(g, imgs, labels, 2.55)
    assert np.trace(arg2) / n >= 0.99, f"argmax confusion with the [0,100] mapping:\n{arg2}"
    assert np.trace(thr2) / n >= 0.96, f"thresholded confusion with the [0,100] mapping:\n{thr2}"


# ---------------------------------------------------------------------------------------------- GPU
def _random_net(rng, c1, c2, nc):
    return dict(conv1_w=rng.normal(0, 0.3, (c1, 3, 3, 1)), conv1_b=rng.normal(0, 0.1, c1),
                conv2_w=rng.normal(0, 0.1, (c2, 3, 3, c1)), conv2_b=rng.normal(0, 0.1, c2),
                bn_scale=rng.uniform(0.5, 1.5, c2), bn_shift=rng.normal(0, 0.2, c2),
                dense_w=rng.normal(0, 0.2, (nc, 2 * c2)), dense_b=rng.normal(0, 0.1, nc))


def _graph_of(net):
    """A tflite_model.Graph of the dam_net architecture around raw tensors (for the numpy checker)."""
    f = {k: np.ascontiguousarray(v, np.float32) for k, v in net.items()}
    t = [T.Tensor("in", (1, 15, 12, 1), np.float32, None)]
    def const(a):
        t.append(T.Tensor("c", a.shape, np.float32, a))
        return len(t) - 1
    def act():
        t.append(T.Tensor("a", (), np.float32, None))
        return len(t) - 1
    conv = {"padding": "VALID", "stride_w": 1, "stride_h": 1, "act": "RELU", "dilation_w": 1, "dilation_h": 1}
    pool = {"padding": "VALID", "stride_w": 2, "stride_h": 2, "filter_w": 2, "filter_h": 2, "act": "NONE"}
    ops, cur = [], 0
    for kind, consts, opt in (("CONV_2D", ("conv1_w", "conv1_b"), conv), ("MAX_POOL_2D", (), pool),
                              ("CONV_2D", ("conv2_w", "conv2_b"), conv), ("MAX_POOL_2D", (), pool),
                              ("MUL", ("bn_scale",), {"act": "NONE"}), ("ADD", ("bn_shift",), {"act": "NONE"}),
                              ("RESHAPE", (), {}), ("FULLY_CONNECTED", ("dense_w", "dense_b"), {"act": "NONE"}),
                              ("SOFTMAX", (), {"beta": 1.0})):
        ins = [cur] + [const(f[c]) for c in consts]
        cur = act()
        ops.append(T.Op(kind, ins, [cur], opt))
    return T.Graph(t, ops, [0], [cur])


@pytest.mark.gpu
def test_gpu_network_matches_numpy_forward_bit_for_bit_on_logits():
    from cones_perception_b200 import api
    g = T.load(model_bytes())
    z = np.load(os.path.join(GOLD, "cone_images.npz"))
    rng = np.random.default_rng(0)
    imgs = np.concatenate([z["real_images"], z["synth_images"], rng.integers(0, 256, (64, 15, 12), dtype=np.uint8),
                           np.zeros((1, 15, 12), np.uint8), np.full((1, 15, 12), 255, np.uint8)])
    with api.ConesGpu(max_points=1024) as h:
        h.color_net_load_tflite(model_bytes())
        colors, probs, logits = h.classify_images(imgs)
        flags_low = 0
        for i, im in enumerate(imgs):
            p, l = D.forward(g, im)
            assert np.array_equal(l.view(np.uint32), logits[i].view(np.uint32)), f"image {i}: logits differ"
            assert np.max(np.abs(p - probs[i])) <= 1e-6, f"image {i}: softmax differs"
            if abs(float(p.max()) - 0.8) > 2e-6:
                assert int(colors[i]) == D.decide(p), f"image {i}"
            else:
                flags_low += 1
        assert flags_low < 5


@pytest.mark.gpu
@pytest.mark.parametrize("c1,c2,nc", [(16, 32, 3), (5, 7, 2), (1, 1, 1), (16, 32, 1), (3, 32, 3)])
def test_gpu_network_random_weights(c1, c2, nc):
    from cones_perception_b200 import api
    rng = np.random.default_rng(c1 * 100 + c2)
    net = _random_net(rng, c1, c2, nc)
    g = _graph_of(net)
    imgs = rng.integers(0, 256, (24, 15, 12), dtype=np.uint8)
    with api.ConesGpu(max_points=1024) as h:
        h.color_net_load(**net)
        colors, probs, logits = h.classify_images(imgs)
        for i, im in enumerate(imgs):
            p, l = D.forward(g, im)
            assert np.array_equal(l.view(np.uint32), logits[i].view(np.uint32)), (i, l, logits[i])
            assert np.max(np.abs(p - probs[i])) <= 1e-6


@pytest.mark.gpu
def test_gpu_refuses_other_models_and_damaged_files():
    from cones_perception_b200 import api
    raw = bytearray(model_bytes())
    with api.ConesGpu(max_points=1024) as h:
        with pytest.raises(api.ConesGpuError) as e:
            h.classify_images(np.zeros((1, 15, 12), np.uint8))
        assert e.value.status == api.CP_E_STATE            # nothing loaded yet
        for bad in (b"", b"TFL3", bytes(64), bytes(raw[:2000]), bytes(raw[:-500])):
            with pytest.raises(api.ConesGpuError) as e:
                h.color_net_load_tflite(bad)
            assert e.value.status == api.CP_E_PARAM
        rng = np.random.default_rng(1)
        refused = 0
        for _ in range(300):                               # random damage: refused or loaded, never a crash
            b = bytearray(raw)
            for pos in rng.integers(0, 4000, 4):           # the head holds the tables and offsets
                b[int(pos)] = int(rng.integers(0, 256))
            try:
                h.color_net_load_tflite(bytes(b))
            except api.ConesGpuError as e:
                assert e.status == api.CP_E_PARAM
                refused += 1
        assert refused > 0
        with pytest.raises(api.ConesGpuError):
            h.color_net_load(**_random_net(rng, 17, 32, 3))
        h.color_net_load_tflite(bytes(raw))
        colors, _, _ = h.classify_images(np.zeros((2, 15, 12), np.uint8))
        assert len(colors) == 2


@pytest.mark.gpu
def test_cone_colors_end_to_end_equals_the_service_chain():
    """cp_cone_colors = get_reconstructed_cone -> to_image -> network -> threshold for every cone of a frame, on
    the cloud cp_detect staged; compared with the same chain on the CPU (oracle crops and raster, numpy network).
    Intensities are scaled into the range the network was trained on so that several colours occur."""
    from cones_perception_b200 import api, scans
    from cones_perception_b200.color_classifier import ColorClassifier
    from cones_perception_b200.pointcloud2 import PointCloud2
    from oracle import oracle as O
    g = T.load(model_bytes())
    cfg = scans.config(2)
    frame = scans.generate(cfg, 1, base_seed=3)[0].copy()
    frame[:, 3] = np.clip(frame[:, 3] * 2.5, 0, 255)
    msg = PointCloud2.from_xyzi(frame)
    with api.ConesGpu(max_points=len(frame)) as h:
        clf = ColorClassifier(h, model_bytes())
        cl, _ = h.detect(msg, cfg.detect, cfg.ground)
        centers = [(float(c["x"]), float(c["y"])) for c in cl] + [(55.0, 55.0)]      # the last box is empty
        colors, probs, flags = h.cone_colors(centers)
        pts = O.points32(frame)
        expect = []
        for k, (cx, cy) in enumerate(centers):
            crop = O.reconstruct_cone(pts, cx, cy)
            if len(crop) == 0:
                assert flags[k] & api.CONE_EMPTY and colors[k] == api.COLOR_NO_ANSWER
                continue
            xyzi = np.stack([crop["x"], crop["y"], crop["z"], crop["intensity"]], 1)
            img, f = O.to_image(xyzi)
            assert f == 0
            p, _ = D.forward(g, img)
            assert np.max(np.abs(p - probs[k])) <= 1e-6, k
            assert int(colors[k]) == D.decide(p), k
            expect.append(D.decide(p))
        assert len(expect) == len(centers) - 1 and len(set(expect)) >= 1
        assert clf.classify(centers) == expect           # the service skips the empty cone: a shorter response
        assert clf.classify([]) == []
