// tflite_reader.hpp — reads the colour classifier out of a TensorFlow-Lite flatbuffer (host code, no TFLite, no
// flatbuffers library).  The reference loads models/dam_net/dam_net.tflite with tf.lite.Interpreter
// (scripts/color_classifier_server.py:66-71, path from the launch parameter ~model_path); the drop-in shell hands
// the same file to cp_color_net_load_tflite.
//
// Only the one architecture cone_color_kernel evaluates is accepted (color_net.cuh): the operator list must be
// exactly CONV_2D, MAX_POOL_2D, CONV_2D, MAX_POOL_2D, MUL, ADD, RESHAPE, FULLY_CONNECTED, SOFTMAX with float32
// tensors, 3x3 VALID stride-1 convolutions with fused ReLU, 2x2 stride-2 VALID pools, softmax beta 1 and a
// 1x15x12x1 input.  Anything else is refused with a message — never evaluated approximately.
// The file is untrusted input: every offset is bounds-checked before it is followed.
//
// Schema slots used (tensorflow/lite/schema/schema.fbs v3): Model{1 operator_codes, 2 subgraphs, 4 buffers},
// OperatorCode{0 deprecated_builtin_code:int8, 1 custom_code, 3 builtin_code:int32}, SubGraph{0 tensors, 1 inputs,
// 2 outputs, 3 operators}, Tensor{0 shape, 1 type, 2 buffer}, Operator{0 opcode_index, 1 inputs, 2 outputs,
// 4 builtin_options}, Buffer{0 data}, Conv2DOptions{0 padding, 1 stride_w, 2 stride_h, 3 activation, 4/5 dilation},
// Pool2DOptions{0 padding, 1 stride_w, 2 stride_h, 3 filter_w, 4 filter_h, 5 activation},
// FullyConnectedOptions{0 activation}, SoftmaxOptions{0 beta}, Mul/AddOptions{0 activation}.
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

namespace cp_tflite {

struct Error {
  std::string what;
};

class Buf {
 public:
  Buf(const uint8_t* p, size_t n) : p_(p), n_(n) {}
  template <typename T>
  T rd(size_t pos) const {
    if (pos > n_ || n_ - pos < sizeof(T)) throw Error{"offset outside the file"};
    T v;
    memcpy(&v, p_ + pos, sizeof(T));
    return v;
  }
  const uint8_t* at(size_t pos, size_t len) const {
    if (pos > n_ || n_ - pos < len) throw Error{"vector outside the file"};
    return p_ + pos;
  }
  size_t size() const { return n_; }

 private:
  const uint8_t* p_;
  size_t n_;
};

struct Table {
  const Buf* b = nullptr;
  size_t pos = 0, vt = 0;
  uint16_t vt_len = 0;
  Table() {}
  Table(const Buf& buf, size_t p) : b(&buf), pos(p) {
    const int32_t so = buf.rd<int32_t>(p);
    const int64_t v = (int64_t)p - so;
    if (v < 0) throw Error{"vtable before the file"};
    vt = (size_t)v;
    vt_len = buf.rd<uint16_t>(vt);
    if (vt_len < 4) throw Error{"short vtable"};
  }
  size_t field(int slot) const {   // absolute position, 0 = absent
    const size_t e = 4 + 2 * (size_t)slot;
    if (e + 2 > vt_len) return 0;
    const uint16_t o = b->rd<uint16_t>(vt + e);
    return o ? pos + o : 0;
  }
  template <typename T>
  T scalar(int slot, T dflt) const {
    const size_t f = field(slot);
    return f ? b->rd<T>(f) : dflt;
  }
  size_t indirect(int slot) const {
    const size_t f = field(slot);
    return f ? f + b->rd<uint32_t>(f) : 0;
  }
  bool has(int slot) const { return field(slot) != 0; }
  Table table(int slot) const {
    const size_t p = indirect(slot);
    if (!p) throw Error{"missing table"};
    return Table(*b, p);
  }
  uint32_t vec_len(int slot) const {
    const size_t p = indirect(slot);
    return p ? b->rd<uint32_t>(p) : 0;
  }
  const uint8_t* vec_data(int slot, size_t elem) const {
    const size_t p = indirect(slot);
    if (!p) return nullptr;
    return b->at(p + 4, (size_t)b->rd<uint32_t>(p) * elem);
  }
  Table vec_table(int slot, uint32_t i) const {
    const size_t p = indirect(slot);
    if (!p || i >= b->rd<uint32_t>(p)) throw Error{"table index out of range"};
    const size_t e = p + 4 + 4 * (size_t)i;
    return Table(*b, e + b->rd<uint32_t>(e));
  }
  int32_t vec_i32(int slot, uint32_t i) const {
    const size_t p = indirect(slot);
    if (!p || i >= b->rd<uint32_t>(p)) throw Error{"index out of range"};
    return b->rd<int32_t>(p + 4 + 4 * (size_t)i);
  }
};

struct ColorNetWeights {
  uint32_t c1 = 0, c2 = 0, classes = 0;
  std::vector<float> conv1_w, conv1_b, conv2_w, conv2_b, bn_scale, bn_shift, dense_w, dense_b;
};

enum { OP_ADD = 0, OP_CONV_2D = 3, OP_FULLY_CONNECTED = 9, OP_MAX_POOL_2D = 17, OP_MUL = 18, OP_RESHAPE = 22, OP_SOFTMAX = 25 };

inline ColorNetWeights read_color_net(const uint8_t* data, size_t bytes) {
  Buf buf(data, bytes);
  if (bytes < 8 || memcmp(data + 4, "TFL3", 4) != 0) throw Error{"not a TFL3 flatbuffer"};
  Table model(buf, buf.rd<uint32_t>(0));
  std::vector<int32_t> codes;
  for (uint32_t i = 0; i < model.vec_len(1); ++i) {
    Table oc = model.vec_table(1, i);
    if (oc.has(1)) throw Error{"custom operators are not supported"};
    int32_t code = oc.scalar<int32_t>(3, 0);
    if (code == 0) code = oc.scalar<int8_t>(0, 0);
    codes.push_back(code);
  }
  if (model.vec_len(2) != 1) throw Error{"expected one subgraph"};
  Table sg = model.vec_table(2, 0);
  const uint32_t n_tensors = sg.vec_len(0), n_buffers = model.vec_len(4);

  auto shape_of = [&](int32_t t) {
    if (t < 0 || (uint32_t)t >= n_tensors) throw Error{"tensor index out of range"};
    Table tt = sg.vec_table(0, (uint32_t)t);
    if (tt.scalar<int8_t>(1, 0) != 0) throw Error{"only float32 tensors are supported"};
    std::vector<int32_t> s(tt.vec_len(0));
    for (uint32_t i = 0; i < s.size(); ++i) s[i] = tt.vec_i32(0, i);
    return s;
  };
  auto constant = [&](int32_t t, size_t expect) {
    if (t < 0 || (uint32_t)t >= n_tensors) throw Error{"tensor index out of range"};
    Table tt = sg.vec_table(0, (uint32_t)t);
    if (tt.scalar<int8_t>(1, 0) != 0) throw Error{"only float32 constants are supported"};
    const uint32_t bi = tt.scalar<uint32_t>(2, 0);
    if (bi >= n_buffers) throw Error{"buffer index out of range"};
    Table bt = model.vec_table(4, bi);
    if (bt.vec_len(0) != expect * 4) throw Error{"constant tensor has an unexpected size"};
    std::vector<float> v(expect);
    if (expect) memcpy(v.data(), bt.vec_data(0, 1), expect * 4);
    return v;
  };
  struct OpView {
    int32_t code;
    std::vector<int32_t> in, out;
    Table opt;
    bool has_opt;
  };
  const int32_t want[9] = {OP_CONV_2D, OP_MAX_POOL_2D, OP_CONV_2D, OP_MAX_POOL_2D, OP_MUL, OP_ADD, OP_RESHAPE, OP_FULLY_CONNECTED, OP_SOFTMAX};
  if (sg.vec_len(3) != 9) throw Error{"unexpected operator count (dam_net has 9)"};
  std::vector<OpView> ops;
  for (uint32_t i = 0; i < 9; ++i) {
    Table o = sg.vec_table(3, i);
    const uint32_t ci = o.scalar<uint32_t>(0, 0);
    if (ci >= codes.size()) throw Error{"opcode index out of range"};
    OpView v;
    v.code = codes[ci];
    if (v.code != want[i]) throw Error{"unexpected operator sequence: not the dam_net architecture"};
    for (uint32_t k = 0; k < o.vec_len(1); ++k) v.in.push_back(o.vec_i32(1, k));
    for (uint32_t k = 0; k < o.vec_len(2); ++k) v.out.push_back(o.vec_i32(2, k));
    v.has_opt = o.has(4);
    if (v.has_opt) v.opt = o.table(4);
    ops.push_back(v);
  }
  // the operators must form a chain: each consumes the previous output
  if (sg.vec_len(1) != 1 || sg.vec_len(2) != 1 || ops[0].in.empty() || ops[0].in[0] != sg.vec_i32(1, 0))
    throw Error{"unexpected graph inputs"};
  for (int i = 1; i < 9; ++i)
    if (ops[i].in.empty() || ops[i - 1].out.size() != 1 || ops[i].in[0] != ops[i - 1].out[0]) throw Error{"operators do not form a chain"};
  if (ops[8].out.size() != 1 || ops[8].out[0] != sg.vec_i32(2, 0)) throw Error{"unexpected graph output"};

  const std::vector<int32_t> in_shape = shape_of(ops[0].in[0]);
  if (in_shape != std::vector<int32_t>{1, 15, 12, 1}) throw Error{"input must be 1x15x12x1"};
  auto check_conv = [&](const OpView& o) {
    if (o.in.size() != 3 || !o.has_opt) throw Error{"CONV_2D needs filter, bias and options"};
    if (o.opt.scalar<int8_t>(0, 0) != 1 /*VALID*/ || o.opt.scalar<int32_t>(1, 0) != 1 || o.opt.scalar<int32_t>(2, 0) != 1 ||
        o.opt.scalar<int8_t>(3, 0) != 1 /*RELU*/ || o.opt.scalar<int32_t>(4, 1) != 1 || o.opt.scalar<int32_t>(5, 1) != 1)
      throw Error{"CONV_2D must be VALID, stride 1, dilation 1, fused ReLU"};
  };
  auto check_pool = [&](const OpView& o) {
    if (!o.has_opt || o.opt.scalar<int8_t>(0, 0) != 1 || o.opt.scalar<int32_t>(1, 0) != 2 || o.opt.scalar<int32_t>(2, 0) != 2 ||
        o.opt.scalar<int32_t>(3, 0) != 2 || o.opt.scalar<int32_t>(4, 0) != 2 || o.opt.scalar<int8_t>(5, 0) != 0)
      throw Error{"MAX_POOL_2D must be 2x2, stride 2, VALID, no activation"};
  };
  ColorNetWeights w;
  check_conv(ops[0]);
  const std::vector<int32_t> f1 = shape_of(ops[0].in[1]);
  if (f1.size() != 4 || f1[1] != 3 || f1[2] != 3 || f1[3] != 1 || f1[0] < 1 || f1[0] > 16) throw Error{"first convolution must be [c1<=16][3][3][1]"};
  w.c1 = (uint32_t)f1[0];
  w.conv1_w = constant(ops[0].in[1], (size_t)w.c1 * 9);
  w.conv1_b = constant(ops[0].in[2], w.c1);
  check_pool(ops[1]);
  check_conv(ops[2]);
  const std::vector<int32_t> f2 = shape_of(ops[2].in[1]);
  if (f2.size() != 4 || f2[1] != 3 || f2[2] != 3 || (uint32_t)f2[3] != w.c1 || f2[0] < 1 || f2[0] > 32) throw Error{"second convolution must be [c2<=32][3][3][c1]"};
  w.c2 = (uint32_t)f2[0];
  w.conv2_w = constant(ops[2].in[1], (size_t)w.c2 * 9 * w.c1);
  w.conv2_b = constant(ops[2].in[2], w.c2);
  check_pool(ops[3]);
  for (int i = 4; i <= 5; ++i) {
    if (ops[i].in.size() != 2) throw Error{"MUL / ADD need one constant operand"};
    if (ops[i].has_opt && ops[i].opt.scalar<int8_t>(0, 0) != 0) throw Error{"MUL / ADD must have no fused activation"};
  }
  w.bn_scale = constant(ops[4].in[1], w.c2);
  w.bn_shift = constant(ops[5].in[1], w.c2);
  if (ops[7].in.size() != 3) throw Error{"FULLY_CONNECTED needs weights and bias"};
  if (ops[7].has_opt && ops[7].opt.scalar<int8_t>(0, 0) != 0) throw Error{"FULLY_CONNECTED must have no fused activation"};
  const std::vector<int32_t> fd = shape_of(ops[7].in[1]);
  if (fd.size() != 2 || (uint32_t)fd[1] != 2 * w.c2 || fd[0] < 1 || fd[0] > 4) throw Error{"dense layer must be [classes<=4][2*c2]"};
  w.classes = (uint32_t)fd[0];
  w.dense_w = constant(ops[7].in[1], (size_t)w.classes * 2 * w.c2);
  w.dense_b = constant(ops[7].in[2], w.classes);
  if (ops[8].has_opt && ops[8].opt.scalar<float>(0, 1.0f) != 1.0f) throw Error{"SOFTMAX beta must be 1"};
  return w;
}

}  // namespace cp_tflite
