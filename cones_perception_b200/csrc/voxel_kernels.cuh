// voxel_kernels.cuh — pcl::VoxelGrid<PointXYZI>::applyFilter restated for batches of frames
// (reference call site: downsample(), src/cone_detection.cpp:240-249; semantics: SURVEY.md
// Appendix A.4).  Stages: per-frame grid setup from the survivors' bounding box -> voxel key
// per survivor (+ the voxel sort's digit histograms) -> (radix sort, radix_sort.cuh) -> segment
// heads -> sequential fp32 mean per voxel in ascending point order (+ voxel offsets per frame).
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace cp {

struct VoxelFrame {
  i32 min_b[3];
  u32 mul1, mul2;    // divb_mul_[1], divb_mul_[2]
  u32 passthrough;   // PCL's "leaf size too small" guard: output = input
  u32 bits;          // ceil(log2(cells))
  u32 pad;
};

struct VoxelK {
  float inv[3];      // inverse_leaf_size_ = 1.0f / (float)leaf
  u32 frame_bits;    // ceil(log2(n_frames))
};

// one thread per frame
__global__ void voxel_setup_kernel(u32 n_frames, VoxelK k, const u32* __restrict__ bbox_key,
                                   const u32* __restrict__ c_off, VoxelFrame* __restrict__ vf, Ctl* ctl) {
  const u32 f = blockIdx.x * blockDim.x + threadIdx.x;
  u32 bits = 0;
  if (f < n_frames) {
    VoxelFrame v;
    v.pad = 0;
    const u32 cnt = c_off[f + 1] - c_off[f];
    if (cnt == 0) {
      v.min_b[0] = v.min_b[1] = v.min_b[2] = 0;
      v.mul1 = v.mul2 = 0;
      v.passthrough = 0;
      v.bits = 0;
    } else {
      float mn[3], mx[3];
      for (int a = 0; a < 3; ++a) {
        mn[a] = ord2f(bbox_key[f * 8 + a]);
        mx[a] = ord2f(bbox_key[f * 8 + 4 + a]);
      }
      long long d[3];
      i32 div_b[3];
      for (int a = 0; a < 3; ++a) {
        d[a] = (long long)__fmul_rn(__fsub_rn(mx[a], mn[a]), k.inv[a]) + 1;
        v.min_b[a] = (i32)floorf(__fmul_rn(mn[a], k.inv[a]));
        const i32 max_b = (i32)floorf(__fmul_rn(mx[a], k.inv[a]));
        div_b[a] = max_b - v.min_b[a] + 1;
      }
      if (d[0] * d[1] * d[2] > 2147483647ll) {
        v.passthrough = 1;
        v.mul1 = v.mul2 = 0;
        v.bits = ceil_log2_u64(cnt);
      } else {
        v.passthrough = 0;
        v.mul1 = (u32)div_b[0];
        v.mul2 = (u32)div_b[0] * (u32)div_b[1];
        const u64 cells = (u64)(u32)div_b[0] * (u32)div_b[1] * (u32)div_b[2];
        v.bits = cells >= (1ull << 31) ? 32u : ceil_log2_u64(cells);
      }
    }
    vf[f] = v;
    bits = v.bits;
  }
  bits = __reduce_max_sync(kFull, bits);
  if (lane_id() == 0 && bits) atomicMax(&ctl->voxel_key_bits, bits);
}

// one thread per survivor: PCL idx (uint32) prefixed by the frame id.  The kernel also fixes the total sort width
// (key bits + frame bits, after voxel_setup) and feeds the voxel sort: digit histograms of every live pass and
// clean look-back words for pass 0 (radix_sort.cuh, sort_feed_*).
__global__ void __launch_bounds__(256) voxel_key_kernel(Ctl* ctl, VoxelK k, const float4* __restrict__ pts,
                                                        const u32* __restrict__ frame,
                                                        const u32* __restrict__ c_off,
                                                        const VoxelFrame* __restrict__ vf, u64* __restrict__ keys,
                                                        u32* __restrict__ vals, u32* sort_hdr, u32* sort_state) {
  __shared__ SortFeedSmem feed;
  const u32 n = ctl->n_surv;
  const u32 kb = ctl->voxel_key_bits;
  const u32 bits = kb + k.frame_bits;
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl->vsort_bits = bits;
  const u32 passes = sort_feed_passes(bits, n);
  sort_feed_begin(feed, passes);
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 base = blockIdx.x * blockDim.x; base < n; base += stride) {   // warp-uniform bounds (sort_feed_key)
    const u32 i = base + threadIdx.x;
    const bool valid = i < n;
    u64 key = 0;
    if (valid) {
      const u32 f = frame[i];
      const VoxelFrame v = vf[f];
      u32 idx;
      if (v.passthrough) {
        idx = i - c_off[f];
      } else {
        const float4 p = pts[i];
        const i32 i0 = (i32)__fsub_rn(floorf(__fmul_rn(p.x, k.inv[0])), (float)v.min_b[0]);
        const i32 i1 = (i32)__fsub_rn(floorf(__fmul_rn(p.y, k.inv[1])), (float)v.min_b[1]);
        const i32 i2 = (i32)__fsub_rn(floorf(__fmul_rn(p.z, k.inv[2])), (float)v.min_b[2]);
        idx = (u32)i0 + (u32)i1 * v.mul1 + (u32)i2 * v.mul2;
      }
      key = ((u64)f << kb) | (u64)idx;
      keys[i] = key;
      vals[i] = i;
    }
    sort_feed_key(feed, key, valid, passes);
  }
  sort_feed_flush(feed, sort_hdr, sort_state, passes, n);
}

// ---- generic "segment heads" pass over a sorted key array ---------------------------------
// head[i] = (i == 0 || key[i] != key[i-1]); writes starts[seg] = i for every head and *d_total = number of
// segments.
constexpr int kHeadThreads = 256;
constexpr int kHeadItems = 4;
constexpr int kHeadTile = kHeadThreads * kHeadItems;

struct HeadArgs {
  const u64* keys_a;
  const u64* keys_b;
  const u32* d_bits;   // selects A or B (result parity of the preceding sort)
  const u32* d_n;
  u32* starts;
  u32 starts_cap;
  u32* d_total;
  u64* desc;
  u32* ticket;
  u32* error;
  u32 err_bit;
  // optional: every head also enters (its key -> its segment number) into an open-addressing hash table
  // (the neighbour grid's cell -> cell id map, cluster_kernels.cuh); hash_mask is read from device memory
  u64* hkeys;
  u32* hvals;
  const u32* d_hash_mask;
};

constexpr u64 kHashEmpty = 0xFFFFFFFFFFFFFFFFull;
__device__ __forceinline__ u32 hash_u64(u64 k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (u32)k;
}
__device__ __forceinline__ void hash_insert_one(u64* hkeys, u32* hvals, u32 mask, u64 key, u32 val) {
  u32 slot = hash_u64(key) & mask;
  for (u32 probe = 0; probe <= mask; ++probe) {
    const u64 old = atomicCAS((unsigned long long*)&hkeys[slot], (unsigned long long)kHashEmpty,
                              (unsigned long long)key);
    if (old == kHashEmpty || old == key) {
      hvals[slot] = val;
      return;
    }
    slot = (slot + 1) & mask;
  }
}

// The look-back of a tile is resolved one tile later (see mask_compact_kernel): a tile publishes its head count,
// its CTA counts the heads of its next tile, and only then comes back for the prefix — which the tiles before it
// have published by then — and writes the parked tile's segment starts.  A parked tile is a few registers.
__global__ void __launch_bounds__(kHeadThreads) segment_heads_kernel(HeadArgs a) {
  const u32 n = *a.d_n;
  const u64* keys = sorted_in_b(*a.d_bits) ? a.keys_b : a.keys_a;
  const u32 tiles = (n + kHeadTile - 1) / kHeadTile;
  __shared__ u32 wsum[kHeadThreads / 32];
  __shared__ u32 s_tile, s_excl;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  bool parked = false;
  u32 p_tile = 0, p_total = 0, p_i0 = 0, p_flags = 0, p_before = 0;   // p_before: heads of the tile before this thread
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    const bool valid = tile < tiles;
    u32 i0 = 0, flags = 0, before = 0, total = 0;
    if (valid) {
      i0 = tile * kHeadTile + threadIdx.x * kHeadItems;
      u64 prev = (i0 > 0 && i0 <= n) ? keys[i0 - 1] : 0ull;
      u32 cnt = 0;
#pragma unroll
      for (int j = 0; j < kHeadItems; ++j) {
        const u32 i = i0 + j;
        if (i < n) {
          const u64 kcur = keys[i];
          if (i == 0 || kcur != prev) {
            flags |= 1u << j;
            ++cnt;
          }
          prev = kcur;
        }
      }
      u32 inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) wsum[warp] = inc;
      __syncthreads();
      u32 woff = 0;
#pragma unroll
      for (int w = 0; w < kHeadThreads / 32; ++w) {
        const u32 t = wsum[w];
        if (w < warp) woff += t;
        total += t;
      }
      before = woff + inc - cnt;
      if (threadIdx.x == 0) lookback_publish(a.desc, tile, total);
    }
    if (parked) {
      if (warp == 0) {
        const u32 e = lookback_resolve(a.desc, p_tile, p_total);
        if (lane == 0) s_excl = e;
      }
      __syncthreads();
      u32 run = s_excl + p_before;
      if (threadIdx.x == 0 && p_tile == tiles - 1) {
        *a.d_total = min(s_excl + p_total, a.starts_cap);
        if (s_excl + p_total > a.starts_cap) atomicOr(a.error, a.err_bit);
      }
#pragma unroll
      for (int j = 0; j < kHeadItems; ++j) {
        const u32 i = p_i0 + j;
        if (i < n) {
          if (p_flags & (1u << j)) {
            if (run < a.starts_cap) {
              a.starts[run] = i;
              if (a.hkeys) hash_insert_one(a.hkeys, a.hvals, *a.d_hash_mask, keys[i], run);
            }
            ++run;
          }
        }
      }
    }
    parked = valid;
    p_tile = tile; p_total = total; p_i0 = i0; p_flags = flags; p_before = before;
    if (!valid) break;
  }
}

// one thread per voxel: CentroidPoint<PointXYZI> = sequential fp32 sums in record order,
// divided by the count (IEEE division)
struct VoxelOut {
  float4* vox;       // [V] x, y, z, intensity means
  u32* vox_frame;    // [V]
  u32* v_off;        // [F+1]
};

// G lanes per voxel: the lanes fetch G records of the segment at a time (index, then point: two dependent loads per
// record, which one thread walking the segment alone would serialise), and every lane then adds the G points in
// record order — the fp32 sums stay sequential, as the reference's accumulator is.  G follows the average segment
// length of the run: a warp per voxel on dense clouds (config 4: 34 points per voxel), two lanes where nearly every
// point is its own voxel (config 5) — with eight lanes throughout, config 4 walked five dependent rounds per voxel
// and config 5 kept six of eight lanes idle.
template <u32 G>
__device__ __forceinline__ void voxel_mean_body(u32 nv, u32 n, u32 kb, const u64* __restrict__ keys,
                                                const u32* __restrict__ vals, const u32* __restrict__ starts,
                                                const float4* __restrict__ pts, const u32* __restrict__ src,
                                                const u32* __restrict__ frame_n, u32 uniform_n,
                                                const u32* __restrict__ gcount, int pad_possible, u32 n_frames,
                                                const VoxelOut& o) {
  const u32 sub = threadIdx.x & (G - 1);
  const u32 gmask = G == 32 ? kFull : (((1u << (G & 31)) - 1u) << (threadIdx.x & 31u & ~(G - 1)));   // this group's lanes
  const u32 ngroups = gridDim.x * blockDim.x / G;
  for (u32 v = (blockIdx.x * blockDim.x + threadIdx.x) / G; v < nv; v += ngroups) {
    const u32 b = starts[v];
    const u32 e = (v + 1 < nv) ? starts[v + 1] : n;
    const u32 f = (u32)(keys[b] >> kb);
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    u32 cnt = e - b;
    for (u32 r0 = b; r0 < e; r0 += G) {
      const u32 r = r0 + sub;
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      bool pad = false;
      if (r < e) {
        const u32 pi = vals[r];
        p = pts[pi];
        pad = pad_possible && src[pi] == 0xFFFFFFFFu;   // (only read when the zero padding can survive the crop)
      }
      const u32 m = e - r0 < G ? e - r0 : G;
      for (u32 k = 0; k < m; ++k) {
        sx = __fadd_rn(sx, __shfl_sync(gmask, p.x, k, G));
        sy = __fadd_rn(sy, __shfl_sync(gmask, p.y, k, G));
        sz = __fadd_rn(sz, __shfl_sync(gmask, p.z, k, G));
        si = __fadd_rn(si, __shfl_sync(gmask, p.w, k, G));
      }
      if (pad_possible && (__ballot_sync(gmask, pad) & gmask)) {
        // the record standing for the ground node's zero padding: adding zeros leaves the
        // sums unchanged, only the count grows by (N - G) - 1
        const u32 nf = uniform_n ? uniform_n : frame_n[f];
        cnt += (nf - gcount[f]) - 1u;
      }
    }
    if (sub == 0) {
      const float c = (float)cnt;
      o.vox[v] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), __fdiv_rn(si, c));
      o.vox_frame[v] = f;
      // v_off[f'] = voxels in frames < f': voxels are sorted by frame, so the first voxel of a frame (and the last
      // voxel of all) settle the offsets of the frames between its predecessor's frame and its own
      const u32 fprev = v ? (u32)(keys[b - 1] >> kb) : 0u;
      if (v == 0)
        for (u32 g = 0; g <= f; ++g) o.v_off[g] = 0;
      else
        for (u32 g = fprev + 1; g <= f; ++g) o.v_off[g] = v;
      if (v == nv - 1)
        for (u32 g = f + 1; g <= n_frames; ++g) o.v_off[g] = nv;
    }
  }
}

__global__ void __launch_bounds__(256) voxel_mean_kernel(const Ctl* __restrict__ ctl, const u64* keys_a,
                                                         const u64* keys_b, const u32* vals_a, const u32* vals_b,
                                                         const u32* __restrict__ starts,
                                                         const float4* __restrict__ pts,
                                                         const u32* __restrict__ src,
                                                         const u32* __restrict__ frame_n, u32 uniform_n,
                                                         const u32* __restrict__ gcount, int pad_possible,
                                                         u32 n_frames, VoxelOut o) {
  const u32 nv = ctl->n_vox, n = ctl->n_surv;
  if (nv == 0 && blockIdx.x == 0)
    for (u32 f = threadIdx.x; f <= n_frames; f += blockDim.x) o.v_off[f] = 0;
  if (nv == 0) return;
  const bool inb = sorted_in_b(ctl->vsort_bits);
  const u64* keys = inb ? keys_b : keys_a;
  const u32* vals = inb ? vals_b : vals_a;
  const u32 kb = ctl->voxel_key_bits;
  const u32 avg = n / nv;   // points per voxel
  if (avg >= 16u)
    voxel_mean_body<32>(nv, n, kb, keys, vals, starts, pts, src, frame_n, uniform_n, gcount, pad_possible, n_frames, o);
  else if (avg >= 3u)
    voxel_mean_body<8>(nv, n, kb, keys, vals, starts, pts, src, frame_n, uniform_n, gcount, pad_possible, n_frames, o);
  else
    voxel_mean_body<2>(nv, n, kb, keys, vals, starts, pts, src, frame_n, uniform_n, gcount, pad_possible, n_frames, o);
}

}  // namespace cp
