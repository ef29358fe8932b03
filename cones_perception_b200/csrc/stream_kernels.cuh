// stream_kernels.cuh — the two streaming passes over the raw scan (16 B/point each):
//
//   ground_sector_min_kernel   pass 1 of GroundRemover::cloud_handler
//                              (src/ground_removal.cpp:58-68): per-sector lowest z
//   mask_crop_compact_kernel   pass 2 (:70-77) fused with filter_points_position
//                              (src/cone_detection.cpp:189-204) and an order-preserving
//                              compaction + bounding box of the survivors
//
// Both are HBM-bound streaming kernels: coalesced 128-bit loads (one float4 per point in
// the compact layout), per-CTA shared-memory reduction, warp ballots for the masks.
// Exactness: every predicate has a cheap fp32 test with a guard band; points that land in
// a guard band (a few per million) re-evaluate the reference's double-precision formula.
#pragma once
#include "common.cuh"

namespace cp {

constexpr int kStreamThreads = 256;
constexpr int kStreamRows = 8;
constexpr int kStreamTile = kStreamThreads * kStreamRows;  // 2048 points

struct Layout {
  u32 step;
  i32 ox, oy, oz, oi;  // byte offsets; oi < 0 => intensity = 0
  u32 mode;            // 0: compact float4 xyzI, 1: 4-byte aligned fields, 2: byte-wise
};

struct Geom {
  u32 n_frames;
  u32 uniform_n;          // > 0: every frame has this many points
  u32 tpf;                // tiles per frame when uniform
  u32 n_tiles;
  const u32* frame_n;     // [F]
  const u64* frame_off;   // [F] first point of the frame (in points)
  const u32* tile_frame;  // [n_tiles] (ragged batches)
  const u32* frame_tile0; // [F] first tile of the frame
};

struct CropK {
  int do_crop;
  float zthr;                // drop iff z < zthr           (== (double)z < level_threshold)
  double smax, smin;         // drop iff s >= smax || s < smin   (s = x^2+y^2+z^2 in double)
  float smax_lo, smax_hi, smin_lo, smin_hi;  // fp32 guard bands around smax / smin
  float f_hi;                // keep iff |atan2f| < f_hi
  float f_lo_guard, f_hi_guard;
};

struct GroundK {
  int do_ground;
  int want_count;   // also count ground survivors per frame (n_ground_kept)
  int pad_survives; // the zero points the ground node pads with survive the crop
};

__device__ __forceinline__ void tile_lookup(const Geom& g, u32 tile, u32& frame, u32& local0, u32& count,
                                            u64& first_point) {
  if (g.uniform_n) {
    frame = tile / g.tpf;
    local0 = (tile - frame * g.tpf) * kStreamTile;
    first_point = (u64)frame * g.uniform_n;
    const u32 n = g.uniform_n;
    count = n - local0 < (u32)kStreamTile ? n - local0 : (u32)kStreamTile;
  } else {
    frame = g.tile_frame[tile];
    local0 = (tile - g.frame_tile0[frame]) * kStreamTile;
    first_point = g.frame_off[frame];
    const u32 n = g.frame_n[frame];
    count = n - local0 < (u32)kStreamTile ? n - local0 : (u32)kStreamTile;
  }
}

__device__ __forceinline__ float load_f32_bytes(const uint8_t* p) {
  u32 v = (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
  return __uint_as_float(v);
}

template <int MODE>
__device__ __forceinline__ float4 load_point(const uint8_t* base, u64 idx, const Layout& L) {
  if (MODE == 0) {
    return ldg_stream(reinterpret_cast<const float4*>(base) + idx);
  } else if (MODE == 1) {
    const uint8_t* p = base + idx * L.step;
    float4 r;
    r.x = __ldg(reinterpret_cast<const float*>(p + L.ox));
    r.y = __ldg(reinterpret_cast<const float*>(p + L.oy));
    r.z = __ldg(reinterpret_cast<const float*>(p + L.oz));
    r.w = L.oi >= 0 ? __ldg(reinterpret_cast<const float*>(p + L.oi)) : 0.0f;
    return r;
  } else {
    const uint8_t* p = base + idx * L.step;
    float4 r;
    r.x = load_f32_bytes(p + L.ox);
    r.y = load_f32_bytes(p + L.oy);
    r.z = load_f32_bytes(p + L.oz);
    r.w = L.oi >= 0 ? load_f32_bytes(p + L.oi) : 0.0f;
    return r;
  }
}

// ---- angle / sector arithmetic --------------------------------------------------------
// Oracle definition of the reference's atan2f: (float)atan2((double)y,(double)x).
__device__ __noinline__ float atan2_exact(float y, float x) {
  return (float)atan2((double)y, (double)x);
}
// src/ground_removal.cpp:20 evaluated at survey time: float((360/16) * M_PI / 180)
#define CP_SECTOR_ANGLE 0.38397244f
#define CP_INV_SECTOR_ANGLE 2.6043537f

// src/ground_removal.cpp:61-64 given the exact float angle
__device__ __forceinline__ int sector_from_exact(float a) {
  const float ang = (a < 0.0f) ? (float)((double)a + 6.283185307179586) : a;
  return (int)floorf(__fdiv_rn(ang, CP_SECTOR_ANGLE));
}

// Fast sector from the fp32 atan2f (<= 2 ulp), exact re-evaluation inside guard bands.
// `a_fast` must be atan2f(y, x).
__device__ __forceinline__ int sector_of(float x, float y, float a_fast) {
  const float ang = a_fast < 0.0f ? a_fast + 6.2831855f : a_fast;
  const float u = ang * CP_INV_SECTOR_ANGLE;
  const float fl = floorf(u);
  const float fr = u - fl;
  const bool risky = (fr < 2e-5f) | (fr > 1.0f - 2e-5f) | (fabsf(a_fast) < 1e-5f);
  if (risky) {
    // common exact case first: +x axis (includes the all-zero filler point)
    if (y == 0.0f && (x > 0.0f || (x == 0.0f && !signbit(x)))) return 0;
    return sector_from_exact(atan2_exact(y, x));
  }
  return (int)fl;
}

// ---- pass 1: per-sector minima ----------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kStreamThreads)
ground_sector_min_kernel(const uint8_t* __restrict__ in, Layout L, Geom g, u32* __restrict__ low_key) {
  __shared__ u32 smin[kSectStride];
  const int lane = lane_id();
  // contiguous chunk of tiles per CTA so the shared table is flushed once per frame change
  const u32 per = (g.n_tiles + gridDim.x - 1) / gridDim.x;
  const u32 t0 = blockIdx.x * per;
  const u32 t1 = t0 + per < g.n_tiles ? t0 + per : g.n_tiles;
  if (threadIdx.x < kSectStride) smin[threadIdx.x] = 0xFFFFFFFFu;
  __syncthreads();
  u32 cur_frame = 0xFFFFFFFFu;
  for (u32 tile = t0; tile < t1; ++tile) {
    u32 frame, local0, count;
    u64 first;
    tile_lookup(g, tile, frame, local0, count, first);
    if (frame != cur_frame) {
      if (cur_frame != 0xFFFFFFFFu) {
        __syncthreads();
        if (threadIdx.x < kNSect && smin[threadIdx.x] != 0xFFFFFFFFu)
          atomicMin(&low_key[cur_frame * kSectStride + threadIdx.x], smin[threadIdx.x]);
        __syncthreads();
        if (threadIdx.x < kSectStride) smin[threadIdx.x] = 0xFFFFFFFFu;
        __syncthreads();
      }
      cur_frame = frame;
    }
    float4 p[kStreamRows];
#pragma unroll
    for (int r = 0; r < kStreamRows; ++r) {
      const u32 i = r * kStreamThreads + threadIdx.x;
      p[r] = (i < count) ? load_point<MODE>(in, first + local0 + i, L) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < kStreamRows; ++r) {
      const u32 i = r * kStreamThreads + threadIdx.x;
      const float x = p[r].x, y = p[r].y, z = p[r].z;
      const bool ok = (i < count) && isfinite(x) && isfinite(y) && isfinite(z);
      int s = 31;  // parking slot for invalid lanes
      u32 zk = 0xFFFFFFFFu;
      if (ok) {
        s = sector_of(x, y, atan2f(y, x));
        zk = f2ord(z);
      }
      // warp-level min per distinct sector (scan-ordered clouds: 1-2 sectors per warp)
      u32 todo = __ballot_sync(kFull, ok);
      while (todo) {
        const int src = __ffs(todo) - 1;
        const int s0 = __shfl_sync(kFull, s, src);
        const bool mine = ok && (s == s0);
        const u32 m = __reduce_min_sync(kFull, mine ? zk : 0xFFFFFFFFu);
        if (lane == src) atomicMin(&smin[s0], m);
        todo &= ~__ballot_sync(kFull, mine);
      }
    }
  }
  __syncthreads();
  if (cur_frame != 0xFFFFFFFFu && threadIdx.x < kNSect && smin[threadIdx.x] != 0xFFFFFFFFu)
    atomicMin(&low_key[cur_frame * kSectStride + threadIdx.x], smin[threadIdx.x]);
}

// ---- pass 2: ground mask + crop + ordered compaction ------------------------------------
struct CompactOut {
  float4* pts;       // [cap] surviving points
  u32* src;          // [cap] frame-local input index
  u32* frame;        // [cap] frame id
  u32 cap;
  u32* c_off;        // [F+1] survivor offsets per frame
  u32* bbox_key;     // [F*8] ordered-int min xyz (0..2) / max xyz (4..6)
  u32* gcount;       // [F] ground survivors (when want_count)
  u64* desc;         // [n_tiles] look-back descriptors
  Ctl* ctl;
  uint8_t* out32;    // node-equivalent output (PCL 32-byte layout) or NULL
};

__device__ __forceinline__ bool crop_keep(const CropK& c, float x, float y, float z, bool& need_angle) {
  need_angle = false;
  if (z < c.zthr) return false;
  const float sf = fmaf(z, z, fmaf(y, y, x * x));
  bool exact = !(sf < 1e30f);
  exact |= (sf > c.smax_lo) & (sf < c.smax_hi);
  exact |= (sf > c.smin_lo) & (sf < c.smin_hi);
  if (exact) {
    // perception_handling::euclidan_dist, src/perception_handling/utils.cpp:32-34
    const double xd = x, yd = y, zd = z;
    double s = __dadd_rn(__dmul_rn(xd, xd), __dmul_rn(yd, yd));
    s = __dadd_rn(s, __dmul_rn(zd, zd));
    if (s >= c.smax || s < c.smin) return false;
  } else {
    if (sf >= c.smax_hi || sf <= c.smin_lo) return false;
  }
  need_angle = true;
  return true;
}

template <int MODE, bool OUT32>
__global__ void __launch_bounds__(kStreamThreads)
mask_crop_compact_kernel(const uint8_t* __restrict__ in, Layout L, Geom g, CropK c, GroundK gk,
                         const u32* __restrict__ low_key, CompactOut o) {
  __shared__ float thr[kSectStride];
  __shared__ u32 wtot[kStreamThreads / 32];
  __shared__ u32 s_tile, s_excl;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  while (true) {
    if (threadIdx.x == 0) s_tile = atomicAdd(&o.ctl->ticket[0], 1u);
    __syncthreads();
    const u32 tile = s_tile;
    if (tile >= g.n_tiles) break;
    u32 frame, local0, count;
    u64 first;
    tile_lookup(g, tile, frame, local0, count, first);
    if (gk.do_ground && threadIdx.x < kNSect) {
      // :75  p.z < low + 0.1 in double  <=>  z < roundup_to_float((double)low + 0.1)
      const double t = (double)ord2f(low_key[frame * kSectStride + threadIdx.x]) + 0.1;
      thr[threadIdx.x] = __double2float_ru(t);
    }
    __syncthreads();

    float4 p[kStreamRows];
    u32 bal[kStreamRows];
    u32 gkept = 0;
    const u32 wbase = warp * (32 * kStreamRows);
#pragma unroll
    for (int r = 0; r < kStreamRows; ++r) {
      const u32 i = wbase + r * 32 + lane;
      p[r] = (i < count) ? load_point<MODE>(in, first + local0 + i, L) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < kStreamRows; ++r) {
      const u32 i = wbase + r * 32 + lane;
      const float x = p[r].x, y = p[r].y, z = p[r].z;
      bool keep = (i < count) && isfinite(x) && isfinite(y) && isfinite(z);
      bool need_angle = false;
      if (keep && c.do_crop) keep = crop_keep(c, x, y, z, need_angle);
      const bool need_sector = gk.do_ground && (keep || gk.want_count) && (i < count) &&
                               isfinite(x) && isfinite(y) && isfinite(z);
      if ((keep && need_angle) || need_sector) {
        const float a_fast = atan2f(y, x);
        if (keep && need_angle) {
          const float aa = fabsf(a_fast);
          if (aa > c.f_lo_guard) {
            if (aa >= c.f_hi_guard) keep = false;
            else keep = fabsf(atan2_exact(y, x)) < c.f_hi;  // src/cone_detection.cpp:200-201
          }
        }
        if (need_sector) {
          const int s = sector_of(x, y, a_fast);
          const bool gkeep = !(z < thr[s]);
          if (gk.want_count && gkeep) gkept++;
          keep = keep && gkeep;
        }
      }
      bal[r] = __ballot_sync(kFull, keep);
    }
    // ---- ranks: warp-contiguous rows => rank order == point order
    u32 wcount = 0;
    u32 rowoff[kStreamRows];
#pragma unroll
    for (int r = 0; r < kStreamRows; ++r) {
      rowoff[r] = wcount;
      wcount += __popc(bal[r]);
    }
    if (lane == 0) wtot[warp] = wcount;
    __syncthreads();
    u32 wexcl = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kStreamThreads / 32; ++w) {
      const u32 t = wtot[w];
      if (w < warp) wexcl += t;
      total += t;
    }
    const bool last_of_frame = (local0 + count) == (g.uniform_n ? g.uniform_n : g.frame_n[frame]);
    const u32 pad_rec = (gk.pad_survives && last_of_frame) ? 1u : 0u;
    if (warp == 0) {
      const u32 e = lookback_exclusive(o.desc, tile, total + pad_rec);
      if (lane == 0) s_excl = e;
    }
    if (gk.want_count) {
      const u32 gsum = __reduce_add_sync(kFull, gkept);
      if (lane == 0 && gsum) atomicAdd(&o.gcount[frame], gsum);
    }
    __syncthreads();
    const u32 excl = s_excl;
    if (threadIdx.x == 0) {
      if (last_of_frame) o.c_off[frame + 1] = excl + total + pad_rec;
      if (tile == g.n_tiles - 1) o.ctl->n_surv = min(excl + total + pad_rec, o.cap);
      if ((u64)excl + total + pad_rec > o.cap) atomicOr(&o.ctl->error, kErrSurvivors);
    }
    if (wcount) {
      u32 mnx = 0xFFFFFFFFu, mny = 0xFFFFFFFFu, mnz = 0xFFFFFFFFu, mxx = 0, mxy = 0, mxz = 0;
#pragma unroll
      for (int r = 0; r < kStreamRows; ++r) {
        if ((bal[r] >> lane) & 1u) {
          const u32 pos = excl + wexcl + rowoff[r] + __popc(bal[r] & lanemask_lt());
          const u32 i = wbase + r * 32 + lane;
          if (pos < o.cap) {
            if (OUT32) {
              float4* dst = reinterpret_cast<float4*>(o.out32 + (u64)pos * 32);
              dst[0] = make_float4(p[r].x, p[r].y, p[r].z, 1.0f);
              dst[1] = make_float4(p[r].w, 0.f, 0.f, 0.f);
            } else {
              o.pts[pos] = p[r];
              o.src[pos] = local0 + i;
              o.frame[pos] = frame;
            }
          }
          const u32 kx = f2ord(p[r].x), ky = f2ord(p[r].y), kz = f2ord(p[r].z);
          mnx = min(mnx, kx); mxx = max(mxx, kx);
          mny = min(mny, ky); mxy = max(mxy, ky);
          mnz = min(mnz, kz); mxz = max(mxz, kz);
        }
      }
      if (!OUT32) {
        mnx = __reduce_min_sync(kFull, mnx); mny = __reduce_min_sync(kFull, mny);
        mnz = __reduce_min_sync(kFull, mnz); mxx = __reduce_max_sync(kFull, mxx);
        mxy = __reduce_max_sync(kFull, mxy); mxz = __reduce_max_sync(kFull, mxz);
        if (lane == 0) {
          u32* bb = o.bbox_key + frame * 8;
          atomicMin(bb + 0, mnx); atomicMin(bb + 1, mny); atomicMin(bb + 2, mnz);
          atomicMax(bb + 4, mxx); atomicMax(bb + 5, mxy); atomicMax(bb + 6, mxz);
        }
      }
    }
    if (!OUT32 && pad_rec && threadIdx.x == 0) {
      // one record stands for the N-G zero points appended by src/ground_removal.cpp:79;
      // its multiplicity is applied when the voxel mean is taken
      const u32 pos = excl + total;
      if (pos < o.cap) {
        o.pts[pos] = make_float4(0.f, 0.f, 0.f, 0.f);
        o.src[pos] = 0xFFFFFFFFu;
        o.frame[pos] = frame;
      }
      const u32 kz = f2ord(0.0f);
      u32* bb = o.bbox_key + frame * 8;
      atomicMin(bb + 0, kz); atomicMin(bb + 1, kz); atomicMin(bb + 2, kz);
      atomicMax(bb + 4, kz); atomicMax(bb + 5, kz); atomicMax(bb + 6, kz);
    }
    __syncthreads();
  }
}

// node-equivalent zero padding, src/ground_removal.cpp:79 (value-initialised PointXYZI)
__global__ void pad_zero_points_kernel(uint8_t* out32, const Ctl* ctl, u32 n) {
  const u32 g0 = ctl->n_surv;
  for (u32 i = g0 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4* dst = reinterpret_cast<float4*>(out32 + (u64)i * 32);
    dst[0] = make_float4(0.f, 0.f, 0.f, 1.0f);
    dst[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

}  // namespace cp
