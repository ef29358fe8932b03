"""TEST INFRASTRUCTURE — numpy forward pass of the reference's colour classifier (`models/dam_net/dam_net.tflite`),
the checker for `cp_cone_colors`.  Only tests/, __graft_entry__.smoke() and bench.py's CPU leg may import this.

Follows what `scripts/color_classifier_server.py:108-120` does with the TFLite interpreter: the 15x12x1 uint8 image
is cast to float32 (no scaling, `:108`), the graph is evaluated, and the answer is `argmax + 1` when the largest
softmax output is >= 0.8, else 0 (`:116-120`).

**Parity unpinned**: TensorFlow-Lite is not installable here, so no interpreter output exists to compare with.
The operators are restated from TFLite's published reference kernels (`tensorflow/lite/kernels/internal/
reference/{conv,pooling,fully_connected,softmax}.h`): NHWC, OHWI filters, VALID padding, accumulation in the order
filter_y, filter_x, in_channel (fp32), bias added after the accumulation, fused ReLU, softmax as
exp(x - max) / sum.  What pins it instead is statistical: on the 577 human-labelled crops the reference ships
(`cones_clouds/cones.pkl`, `color` column) the decision must agree with the label at the rate asserted in
`tests/test_dam_net.py`.
"""
from __future__ import annotations

import numpy as np

COLORS = (None, "yellow", "blue", "orange")     # scripts/color_classifier_server.py:74
THRESHOLD = 0.8                                 # :116


def _act(x, act):
    if act == "NONE":
        return x
    if act == "RELU":
        return np.maximum(x, np.float32(0))
    raise ValueError(act)


def conv2d_valid(x, w, b, act):
    """x [H,W,Ci], w [Co,kh,kw,Ci], sequential fp32 accumulation over (ky, kx, ci) like the reference kernel."""
    H, W, Ci = x.shape
    Co, kh, kw, _ = w.shape
    Ho, Wo = H - kh + 1, W - kw + 1
    acc = np.zeros((Ho, Wo, Co), np.float32)
    for ky in range(kh):
        for kx in range(kw):
            for ci in range(Ci):
                acc += x[ky:ky + Ho, kx:kx + Wo, ci, None] * w[None, None, :, ky, kx, ci]
    return _act(acc + b[None, None, :], act)


def maxpool_valid(x, fh, fw, sh, sw):
    H, W, C = x.shape
    Ho, Wo = (H - fh) // sh + 1, (W - fw) // sw + 1
    out = np.full((Ho, Wo, C), -np.inf, np.float32)
    for dy in range(fh):
        for dx in range(fw):
            out = np.maximum(out, x[dy:dy + sh * Ho:sh, dx:dx + sw * Wo:sw, :])
    return out


def forward(graph, image_u8: np.ndarray):
    """One 15x12 uint8 image -> (softmax probabilities [3], logits [3]), all arithmetic fp32."""
    t = {}
    inp = graph.tensors[graph.inputs[0]]
    t[graph.inputs[0]] = image_u8.astype(np.float32).reshape(inp.shape[1:])
    const = lambda i: graph.tensors[i].data
    logits = None
    for op in graph.ops:
        x = t[op.inputs[0]]
        o = op.options
        if op.kind == "CONV_2D":
            assert o["padding"] == "VALID" and o["stride_w"] == o["stride_h"] == 1
            y = conv2d_valid(x, const(op.inputs[1]), const(op.inputs[2]), o["act"])
        elif op.kind == "MAX_POOL_2D":
            assert o["padding"] == "VALID" and o["act"] == "NONE"
            y = maxpool_valid(x, o["filter_h"], o["filter_w"], o["stride_h"], o["stride_w"])
        elif op.kind == "MUL":
            y = _act(x * const(op.inputs[1]), o["act"])
        elif op.kind == "ADD":
            y = _act(x + const(op.inputs[1]), o["act"])
        elif op.kind == "RESHAPE":
            y = x.reshape(-1)                       # NHWC row-major flatten
        elif op.kind == "FULLY_CONNECTED":
            w, b = const(op.inputs[1]), const(op.inputs[2])
            acc = np.zeros(w.shape[0], np.float32)
            for k in range(w.shape[1]):             # sequential over the input features
                acc += x[k] * w[:, k]
            y = _act(acc + b, o["act"])
            logits = y
        elif op.kind == "SOFTMAX":
            z = (x - x.max()) * np.float32(o["beta"])
            e = np.exp(z.astype(np.float32)).astype(np.float32)
            s = np.float32(0)
            for v in e:
                s = np.float32(s + v)
            y = (e / s).astype(np.float32)
        else:
            raise ValueError(op.kind)
        t[op.outputs[0]] = y.astype(np.float32)
    return t[graph.outputs[0]], logits


def decide(probs) -> int:
    """scripts/color_classifier_server.py:116-120: 0 unknown, 1 yellow, 2 blue, 3 orange."""
    return int(np.argmax(probs)) + 1 if float(np.max(probs)) >= THRESHOLD else 0
