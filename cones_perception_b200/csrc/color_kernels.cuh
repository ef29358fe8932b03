// color_kernels.cuh — inputs of the colour path: what the reference computes
// between a detected cone and the classifier network.
//
//   cone_box_mask_kernel     get_reconstructed_cone (src/cone_detection.cpp:222-238) for up to
//                            64 cone centres in ONE pass over the raw cloud: per point a 64-bit
//                            set of matching boxes, per (centre, 32-point row) a ballot word
//   cone_box_scan_kernel     exclusive scan of the per-(centre, tile) counts -> crop offsets
//   cone_box_gather_kernel   order-preserving gather of the matches into packed crops
//   cone_raster_kernel       ColorClassifier.to_image (scripts/color_classifier_server.py:
//                            130-156): 15x12 range image of a crop, one CTA per cone
//
// The reference walks the whole cloud once per cone on the host and ships each crop through a
// ROS service; here the cloud already sits in HBM from the detection pass and only 180 bytes
// per cone travel back.
//
// Exactness.  The box test is the reference's double-precision comparison folded into four
// fp32 bounds on the host (largest / smallest float on the right side of each double bound),
// so the crops are bit-exact.  The raster evaluates the numpy formulas in fp64; the only
// operation that is not correctly rounded on the device is atan2 (<= 2 ulp), which matters
// only if a pixel coordinate lands within a guard band of a rounding boundary (x.5): such a
// cone is flagged CP_CONE_AMBIGUOUS instead of being silently different.
#pragma once
#include "stream_kernels.cuh"

namespace cp {

constexpr int kConeChunk = 64;  // centres per launch: one bit each in the per-point match set
constexpr int kImgRows = 15, kImgCols = 12, kImgPix = kImgRows * kImgCols;
constexpr u32 kConeEmpty = 1u, kConeBadIndex = 2u, kConeBadIntensity = 4u, kConeAmbiguous = 8u;

struct ConeBoxes {
  u32 n;
  float xlo[kConeChunk], xhi[kConeChunk], ylo[kConeChunk], yhi[kConeChunk];
};

// mask[k][row] = ballot of "point row*32+lane lies in box k"; tile_count[k][tile] = matches of box k
// in the 2048-point tile.  Both start from zero (only non-empty rows are written).
template <int MODE>
__global__ void __launch_bounds__(kStreamThreads) cone_box_mask_kernel(
    const uint8_t* __restrict__ in, Layout L, u64 first_point, u32 n_points, const __grid_constant__ ConeBoxes B,
    u32 n_rows, u32 n_tiles, u32* __restrict__ mask, u32* __restrict__ tile_count) {
  // one 32-point row per warp (a single frame is small: many short CTAs beat 64 long ones)
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 row = blockIdx.x * kStreamWarps + warp;
  if (row >= n_rows) return;
  const u32 tile = row / kTileWords;
  const u32 idx = row * 32 + lane;
  const bool valid = idx < n_points;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) p = load_point<MODE>(in, first_point + idx, L);
  u32 m0 = 0, m1 = 0;
  for (u32 k = 0; k < B.n; ++k) {
    // NaN coordinates fail every comparison, like the reference's double comparisons
    const bool inside = valid && p.x >= B.xlo[k] && p.x <= B.xhi[k] && p.y >= B.ylo[k] && p.y <= B.yhi[k];
    if (k < 32) m0 |= (u32)inside << k;
    else m1 |= (u32)inside << (k - 32);
  }
  u32 any0 = __reduce_or_sync(0xFFFFFFFFu, m0), any1 = __reduce_or_sync(0xFFFFFFFFu, m1);
  while (any0 | any1) {
    const bool hi = any0 == 0;
    const u32 bit = hi ? __ffs(any1) - 1 : __ffs(any0) - 1;
    const u32 b = __ballot_sync(0xFFFFFFFFu, ((hi ? m1 : m0) >> bit) & 1u);
    const u32 k = bit + (hi ? 32u : 0u);
    if (lane == 0) {
      mask[(size_t)k * n_rows + row] = b;
      atomicAdd(&tile_count[(size_t)k * n_tiles + tile], (u32)__popc(b));
    }
    if (hi) any1 &= any1 - 1;
    else any0 &= any0 - 1;
  }
}

// One CTA.  Warp w scans the tile counts of centres w, w+32 of this chunk; thread 0 then chains the
// chunk totals onto crop_off (crop_off[k0] was written by the previous chunk, 0 for the first).
__global__ void __launch_bounds__(1024) cone_box_scan_kernel(u32 n_centers, u32 k0, u32 n_tiles,
                                                             const u32* __restrict__ tile_count,
                                                             u32* __restrict__ tile_excl, u32* __restrict__ crop_off) {
  __shared__ u32 total[kConeChunk];
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (u32 k = warp; k < n_centers; k += 32) {
    u32 carry = 0;
    for (u32 base = 0; base < n_tiles; base += 32) {
      const u32 t = base + lane;
      const u32 c = t < n_tiles ? tile_count[(size_t)k * n_tiles + t] : 0u;
      u32 incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const u32 o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += o;
      }
      if (t < n_tiles) tile_excl[(size_t)k * n_tiles + t] = carry + incl - c;
      carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) total[k] = carry;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    u64 run = crop_off[k0];
    for (u32 k = 0; k < n_centers; ++k) {
      run += total[k];
      crop_off[k0 + k + 1] = run > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)run;  // saturates: the host reports capacity
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kStreamThreads) cone_box_gather_kernel(
    const uint8_t* __restrict__ in, Layout L, u64 first_point, u32 n_centers, u32 k0, u32 n_rows, u32 n_tiles,
    const u32* __restrict__ mask, const u32* __restrict__ tile_count, const u32* __restrict__ tile_excl,
    const u32* __restrict__ crop_off, u32 cap, float4* __restrict__ crop_pts) {
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 tile = blockIdx.x;
  for (u32 k = 0; k < n_centers; ++k) {
    if (tile_count[(size_t)k * n_tiles + tile] == 0) continue;  // uniform over the CTA
    const u32* mrow = mask + (size_t)k * n_rows + (size_t)tile * kTileWords;
    const u32 r0 = tile * kTileWords + lane, r1 = r0 + 32;
    const u32 w0 = r0 < n_rows ? mrow[lane] : 0u, w1 = r1 < n_rows ? mrow[lane + 32] : 0u;
    const u32 c0 = __popc(w0), c1 = __popc(w1);
    u32 s0 = c0, s1 = c1;  // inclusive scans over the 64 rows of the tile
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u32 a = __shfl_up_sync(0xFFFFFFFFu, s0, d), b = __shfl_up_sync(0xFFFFFFFFu, s1, d);
      if ((int)lane >= d) { s0 += a; s1 += b; }
    }
    s1 += __shfl_sync(0xFFFFFFFFu, s0, 31);
    const u64 base = (u64)crop_off[k0 + k] + tile_excl[(size_t)k * n_tiles + tile];
#pragma unroll
    for (int r = 0; r < kStreamRows; ++r) {
      const u32 rr = warp * kStreamRows + r;  // row inside the tile: warps own consecutive rows
      const u32 word = rr < 32 ? __shfl_sync(0xFFFFFFFFu, w0, rr) : __shfl_sync(0xFFFFFFFFu, w1, rr - 32);
      const u32 excl = rr < 32 ? __shfl_sync(0xFFFFFFFFu, s0 - c0, rr) : __shfl_sync(0xFFFFFFFFu, s1 - c1, rr - 32);
      if ((word >> lane) & 1u) {
        const u64 dst = base + excl + __popc(word & ((1u << lane) - 1u));
        if (dst < cap) {
          const u32 idx = (tile * kTileWords + rr) * 32 + lane;
          crop_pts[dst] = load_point<MODE>(in, first_point + idx, L);
        }
      }
    }
  }
}

// ---- range image ---------------------------------------------------------------------
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v = fmin(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v = fmax(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
  return v;
}

// distance of t from the nearest rounding boundary of np.round (k + 0.5)
__device__ __forceinline__ double half_distance(double t) { return fabs(fabs(t - floor(t)) - 0.5); }

constexpr int kRasterThreads = 128;
constexpr double kRad2Deg = 57.29577951308232;  // 180.0 / np.pi (np.degrees multiplies by this constant)

__device__ __forceinline__ double horiz_deg(float4 p) { return __dmul_rn(atan2((double)p.y, (double)p.x), kRad2Deg); }
__device__ __forceinline__ double vert_deg(float4 p) {
  const double X = p.x, Y = p.y;
  const double s = __dadd_rn(__dmul_rn(X, X), __dmul_rn(Y, Y));  // pow(X, 2.0) + pow(Y, 2.0): numpy squares
  return __dmul_rn(atan2((double)p.z, __dsqrt_rn(s)), kRad2Deg);
}

// grid = cones; crops packed as float4 {x, y, z, intensity}, cone c = [off[c], off[c+1])
__global__ void __launch_bounds__(kRasterThreads) cone_raster_kernel(const float4* __restrict__ pts,
                                                                     const u32* __restrict__ off, u32 cap,
                                                                     uint8_t* __restrict__ images,
                                                                     u32* __restrict__ flags_out) {
  __shared__ double s_min[kRasterThreads / 32], s_max[kRasterThreads / 32];
  __shared__ int s_last[kImgPix];
  __shared__ u32 s_flags;
  const u32 cone = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 b = off[cone], e = min(off[cone + 1], cap);
  const u32 n = e > b ? e - b : 0u;
  for (u32 i = tid; i < kImgPix; i += kRasterThreads) s_last[i] = -1;
  if (tid == 0) s_flags = n ? 0u : kConeEmpty;  // color_classifier_server.py:83-84 skips empty clouds
  __syncthreads();
  u32 flags = 0;
  double hmin = __longlong_as_double(0x7ff0000000000000LL), hmax = -hmin;
  bool same_xy = true;  // every point at the first point's (x, y): a zero horizontal range is then exact
  for (u32 i = tid; i < n; i += kRasterThreads) {
    const float4 p = pts[b + i];
    const double ha = horiz_deg(p), va = vert_deg(p);
    same_xy = same_xy && p.x == pts[b].x && p.y == pts[b].y;
    if (!isfinite(ha) || !isfinite(va)) flags |= kConeBadIndex;
    if (!(p.w >= 0.0f && p.w <= 255.0f)) flags |= kConeBadIntensity;  // interp1d([0, 255], ...) bounds
    hmin = fmin(hmin, ha);
    hmax = fmax(hmax, ha);
  }
  hmin = warp_min(hmin);
  hmax = warp_max(hmax);
  if (lane == 0) { s_min[warp] = hmin; s_max[warp] = hmax; }
  if (flags) atomicOr(&s_flags, flags);
  const int all_same = __syncthreads_and(same_xy);
  hmin = s_min[0]; hmax = s_max[0];
#pragma unroll
  for (int w = 1; w < kRasterThreads / 32; ++w) { hmin = fmin(hmin, s_min[w]); hmax = fmax(hmax, s_max[w]); }
  const bool usable = s_flags == 0;
  __syncthreads();
  if (usable) {
    // slope_horiz = (IMG_COLS - 1) / (max - min + 1e-16)                                   (:148)
    const double slope_h = __ddiv_rn(11.0, __dadd_rn(__dsub_rn(hmax, hmin), 1e-16));
    // a 2-ulp atan2 error moves an angle by < 4e-13 degrees; scaled by the slope it bounds the pixel error
    const double guard_h = 1e-9 + slope_h * 4e-13, guard_v = 1e-9;
    // distinct points whose azimuths collapse to one value here need not collapse in libm
    flags = (hmax == hmin && !all_same) ? kConeAmbiguous : 0u;
    for (u32 i = tid; i < n; i += kRasterThreads) {
      const float4 p = pts[b + i];
      const double ha = horiz_deg(p), va = vert_deg(p);
      // slope_vert * (vert_angles - MIN_V_ANGLE), slope_vert = 15 / (-15 - 15)               (:139)
      const double tv = __dmul_rn(-0.5, __dsub_rn(va, -15.0));
      const double dh = __dsub_rn(ha, hmin);
      const double th = __dmul_rn(slope_h, dh);
      const double v = rint(tv), hh = rint(th);  // np.round: half to even
      if (half_distance(tv) < guard_v) flags |= kConeAmbiguous;
      if (dh != 0.0 && half_distance(th) < guard_h) flags |= kConeAmbiguous;
      if (!(v >= -kImgRows && v <= kImgRows - 1) || !(hh >= -kImgCols && hh <= kImgCols - 1)) {
        flags |= kConeBadIndex;  // numpy raises IndexError
        continue;
      }
      const int row = v < 0 ? (int)v + kImgRows : (int)v;  // negative numpy indices wrap
      const int col = hh < 0 ? (int)hh + kImgCols : (int)hh;
      atomicMax(&s_last[row * kImgCols + col], (int)i);  // repeated pixels: the last point wins
    }
    if (flags) atomicOr(&s_flags, flags);
  }
  __syncthreads();
  const bool draw = (s_flags & ~kConeAmbiguous) == 0;
  for (u32 i = tid; i < kImgPix; i += kRasterThreads) {
    const int src = s_last[i];
    uint8_t val = 0;
    if (draw && src >= 0) val = (uint8_t)(int)pts[b + src].w;  // identity interp1d, uint8 store truncates
    images[(size_t)cone * kImgPix + i] = val;
  }
  if (tid == 0) flags_out[cone] = s_flags;
}

}  // namespace cp
