"""A minimal reader for TensorFlow-Lite flatbuffers — just enough to load the colour classifier the reference
ships (`models/dam_net/dam_net.tflite`, loaded by `scripts/color_classifier_server.py:66-71` through
`tf.lite.Interpreter`).  No TensorFlow, no `flatbuffers` package: the flatbuffer wire format is walked by hand.

Only what a small float32 convolutional classifier needs is understood (CONV_2D, MAX_POOL_2D, RESHAPE, MUL, ADD,
FULLY_CONNECTED, SOFTMAX, float32 tensors); anything else raises `UnsupportedModel`, so that a different model is
refused instead of being evaluated wrongly.

Schema facts used (tensorflow/lite/schema/schema.fbs, v3):
  Model        : 0 version, 1 operator_codes, 2 subgraphs, 3 description, 4 buffers
  OperatorCode : 0 deprecated_builtin_code (int8), 1 custom_code, 2 version, 3 builtin_code (int32)
  SubGraph     : 0 tensors, 1 inputs, 2 outputs, 3 operators, 4 name
  Tensor       : 0 shape, 1 type, 2 buffer, 3 name, 4 quantization
  Operator     : 0 opcode_index, 1 inputs, 2 outputs, 3 builtin_options_type, 4 builtin_options
  Buffer       : 0 data
  Conv2DOptions        : 0 padding, 1 stride_w, 2 stride_h, 3 fused_activation, 4 dilation_w, 5 dilation_h
  Pool2DOptions        : 0 padding, 1 stride_w, 2 stride_h, 3 filter_width, 4 filter_height, 5 fused_activation
  FullyConnectedOptions: 0 fused_activation, 1 weights_format, 2 keep_num_dims
  SoftmaxOptions       : 0 beta          Mul/AddOptions: 0 fused_activation
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np


class UnsupportedModel(ValueError):
    pass


BUILTIN = {0: "ADD", 1: "AVERAGE_POOL_2D", 3: "CONV_2D", 4: "DEPTHWISE_CONV_2D", 9: "FULLY_CONNECTED",
           17: "MAX_POOL_2D", 18: "MUL", 19: "RELU", 22: "RESHAPE", 25: "SOFTMAX", 14: "LOGISTIC"}
ACT = {0: "NONE", 1: "RELU", 2: "RELU_N1_TO_1", 3: "RELU6", 4: "TANH"}
PADDING = {0: "SAME", 1: "VALID"}
TENSOR_TYPE = {0: np.float32, 1: np.float16, 2: np.int32, 3: np.uint8, 4: np.int64, 9: np.int8}


class _Table:
    """One flatbuffer table: field slot -> absolute position (0 when the field is absent)."""

    def __init__(self, buf: bytes, pos: int):
        self.buf, self.pos = buf, pos
        vt = pos - struct.unpack_from("<i", buf, pos)[0]
        vt_len = struct.unpack_from("<H", buf, vt)[0]
        self.slots = [struct.unpack_from("<H", buf, vt + 4 + 2 * i)[0] for i in range((vt_len - 4) // 2)]

    def _at(self, slot: int) -> int:
        return self.pos + self.slots[slot] if slot < len(self.slots) and self.slots[slot] else 0

    def scalar(self, slot: int, fmt: str, default=0):
        p = self._at(slot)
        return struct.unpack_from("<" + fmt, self.buf, p)[0] if p else default

    def _indirect(self, slot: int) -> int:
        p = self._at(slot)
        return p + struct.unpack_from("<I", self.buf, p)[0] if p else 0

    def table(self, slot: int):
        p = self._indirect(slot)
        return _Table(self.buf, p) if p else None

    def string(self, slot: int) -> str:
        p = self._indirect(slot)
        if not p:
            return ""
        n = struct.unpack_from("<I", self.buf, p)[0]
        return self.buf[p + 4:p + 4 + n].decode("utf-8", "replace")

    def vector(self, slot: int, dtype) -> np.ndarray:
        p = self._indirect(slot)
        if not p:
            return np.zeros(0, dtype)
        n = struct.unpack_from("<I", self.buf, p)[0]
        return np.frombuffer(self.buf, dtype=dtype, count=n, offset=p + 4).copy()

    def tables(self, slot: int) -> list["_Table"]:
        p = self._indirect(slot)
        if not p:
            return []
        n = struct.unpack_from("<I", self.buf, p)[0]
        out = []
        for i in range(n):
            e = p + 4 + 4 * i
            out.append(_Table(self.buf, e + struct.unpack_from("<I", self.buf, e)[0]))
        return out


@dataclass
class Tensor:
    name: str
    shape: tuple
    dtype: type
    data: np.ndarray | None      # constant tensors only


@dataclass
class Op:
    kind: str
    inputs: list
    outputs: list
    options: dict = field(default_factory=dict)


@dataclass
class Graph:
    tensors: list
    ops: list
    inputs: list
    outputs: list
    description: str = ""


def load(path_or_bytes) -> Graph:
    buf = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    buf = bytes(buf)
    if len(buf) < 8 or buf[4:8] != b"TFL3":
        raise UnsupportedModel("not a TFL3 flatbuffer")
    model = _Table(buf, struct.unpack_from("<I", buf, 0)[0])
    codes = []
    for oc in model.tables(1):
        code = oc.scalar(3, "i", 0) or oc.scalar(0, "b", 0)
        if oc.string(1):
            raise UnsupportedModel(f"custom operator {oc.string(1)!r}")
        codes.append(code)
    buffers = [b.vector(0, np.uint8) for b in model.tables(4)]
    sgs = model.tables(2)
    if len(sgs) != 1:
        raise UnsupportedModel(f"{len(sgs)} subgraphs")
    sg = sgs[0]
    tensors = []
    for t in sg.tables(0):
        ttype = t.scalar(1, "b", 0)
        if ttype not in TENSOR_TYPE:
            raise UnsupportedModel(f"tensor type {ttype}")
        dt = TENSOR_TYPE[ttype]
        shape = tuple(int(v) for v in t.vector(0, np.int32))
        raw = buffers[t.scalar(2, "I", 0)]
        data = None
        if raw.size:
            data = raw.view(dt).reshape(shape) if shape else raw.view(dt)
        tensors.append(Tensor(t.string(3), shape, dt, data))
    ops = []
    for o in sg.tables(3):
        code = codes[o.scalar(0, "I", 0)]
        kind = BUILTIN.get(code)
        if kind is None:
            raise UnsupportedModel(f"builtin operator code {code}")
        opt = o.table(4)
        options = {}
        if kind == "CONV_2D" and opt is not None:
            options = {"padding": PADDING[opt.scalar(0, "b")], "stride_w": opt.scalar(1, "i"),
                       "stride_h": opt.scalar(2, "i"), "act": ACT[opt.scalar(3, "b")],
                       "dilation_w": opt.scalar(4, "i", 1), "dilation_h": opt.scalar(5, "i", 1)}
        elif kind in ("MAX_POOL_2D", "AVERAGE_POOL_2D") and opt is not None:
            options = {"padding": PADDING[opt.scalar(0, "b")], "stride_w": opt.scalar(1, "i"),
                       "stride_h": opt.scalar(2, "i"), "filter_w": opt.scalar(3, "i"),
                       "filter_h": opt.scalar(4, "i"), "act": ACT[opt.scalar(5, "b")]}
        elif kind == "FULLY_CONNECTED":
            options = {"act": ACT[opt.scalar(0, "b")] if opt is not None else "NONE",
                       "keep_num_dims": bool(opt.scalar(2, "b")) if opt is not None else False}
        elif kind == "SOFTMAX":
            options = {"beta": opt.scalar(0, "f", 1.0) if opt is not None else 1.0}
        elif kind in ("MUL", "ADD"):
            options = {"act": ACT[opt.scalar(0, "b")] if opt is not None else "NONE"}
        ops.append(Op(kind, [int(v) for v in o.vector(1, np.int32)], [int(v) for v in o.vector(2, np.int32)], options))
    return Graph(tensors, ops, [int(v) for v in sg.vector(1, np.int32)], [int(v) for v in sg.vector(2, np.int32)],
                 model.string(3))
