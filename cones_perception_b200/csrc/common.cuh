// common.cuh — shared device helpers for libconesgpu (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cp {

typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;

constexpr int kNSect = 17;        // floor(float(2*pi) / 0.38397244f) + 1, SURVEY Appendix C Q2
constexpr int kSectStride = 32;   // per-frame stride of the sector tables
constexpr u32 kFull = 0xFFFFFFFFu;

// ---- device control block: every data-dependent size lives here, so the host never
// ---- has to synchronise between stages (the whole pipeline is one stream of launches).
struct Ctl {
  u32 n_surv;          // C_total: survivors of ground + crop
  u32 n_vox;           // V_total
  u32 n_cells;         // occupied neighbour-grid cells
  u32 n_comp;          // connected components
  u32 n_clusters;      // K_total
  u32 voxel_key_bits;  // max over frames of ceil(log2(cells))
  u32 vsort_bits;      // voxel sort:   voxel_key_bits + frame bits
  u32 csort_bits;      // cell sort
  u32 lsort_bits;      // label sort
  u32 osort_bits;      // cluster-order sort
  u32 hash_mask;       // live hash capacity - 1
  u32 error;           // sticky capacity / internal error bits
  u32 ticket[4];       // dynamic tile tickets
  u32 fast_overflow;   // frames the shared-memory back half could not hold
  u32 fast_max_c;      // largest per-frame survivor count seen by it
  u32 fast_max_v;      // largest per-frame voxel count seen by it
  u32 rows_loaded;     // 32-point rows the keep-mask pass actually read (the rest were skipped as all-ground)
  u32 pad_;
  unsigned long long pairs_visited;  // clustering: candidate voxel pairs looked at (incl. the one-load exits)
  unsigned long long pairs_tested;   // ... of which the squared distance was evaluated against r2
};

enum : u32 { kErrSurvivors = 1u, kErrVoxels = 2u, kErrHash = 4u, kErrInternal = 8u, kErrGather = 16u };

// monotone float <-> uint mapping for atomic min/max on floats
__host__ __device__ __forceinline__ u32 f2ord(float f) {
#ifdef __CUDA_ARCH__
  u32 b = __float_as_uint(f);
#else
  union { float f; u32 u; } c; c.f = f; u32 b = c.u;
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(u32 k) {
  u32 b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; u32 u; } c; c.u = b; return c.f;
#endif
}

__device__ __forceinline__ u32 lanemask_lt() {
  u32 m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// streaming 128-bit load that does not allocate in L1 (input is read once per pass)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ u32 ceil_log2_u64(u64 v) {  // smallest b with 2^b >= v
  if (v <= 1) return 0;
  return 64u - (u32)__clzll((long long)(v - 1));
}

// ---- decoupled look-back over packed 64-bit tile descriptors -------------------------
// descriptor = status (bits 62..63: 0 empty, 1 aggregate, 2 inclusive prefix) | value (low 32)
constexpr u64 kStAgg = 1ull << 62, kStInc = 2ull << 62;

__device__ __forceinline__ void desc_store(u64* p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 desc_load(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// The look-back in two steps, so that a kernel can publish a tile's aggregate as soon as it is known and come back
// for the prefix later (by then the tiles before it have long published theirs and nothing spins):
//   lookback_publish   one thread, as early as possible;
//   lookback_resolve   ONE full warp of the CTA that owns `tile`: returns the exclusive prefix of `tile` and
//                      publishes its inclusive prefix.
// Tiles are handed out by ticket and a CTA publishes a tile's aggregate before it waits for anything, so every
// predecessor a resolve spins on is published by a running CTA: the spin cannot deadlock.
__device__ __forceinline__ void lookback_publish(u64* desc, u32 tile, u32 aggregate) {
  desc_store(desc + tile, (tile == 0 ? kStInc : kStAgg) | aggregate);
}
__device__ __forceinline__ u32 lookback_resolve(u64* desc, u32 tile, u32 aggregate) {
  const int lane = lane_id();
  if (tile == 0) return 0;
  u32 excl = 0;
  i32 base = (i32)tile - 1;
  // kLookWin x 32 predecessors per round trip: with several hundred tiles in flight the nearest inclusive prefix
  // can lie that many tiles back
  constexpr int kLookWin = 4;
  bool done = false;
  while (!done) {
    u64 d[kLookWin];
#pragma unroll
    for (int w = 0; w < kLookWin; ++w) {   // all loads in flight together
      const i32 t = base - w * 32 - lane;
      d[w] = t >= 0 ? desc_load(desc + t) : kStInc;  // lanes before tile 0 read as "inclusive 0"
    }
#pragma unroll
    for (int w = 0; w < kLookWin; ++w) {
      if (done) continue;                  // uniform across the warp
      const i32 t = base - w * 32 - lane;
      while ((d[w] >> 62) == 0) d[w] = desc_load(desc + t);   // not published yet
      const u32 inc_mask = __ballot_sync(kFull, (d[w] >> 62) == 2);
      u32 v = (u32)d[w];
      if (inc_mask) {
        const int first = __ffs(inc_mask) - 1;  // nearest predecessor with an inclusive prefix
        if (lane > first) v = 0;
        done = true;
      }
      excl += __reduce_add_sync(kFull, v);
    }
    base -= 32 * kLookWin;
  }
  if (lane == 0) desc_store(desc + tile, kStInc | (u64)(excl + aggregate));
  return excl;
}
// both steps at once (called by ONE full warp)
__device__ __forceinline__ u32 lookback_exclusive(u64* desc, u32 tile, u32 aggregate) {
  if (lane_id() == 0) lookback_publish(desc, tile, aggregate);
  __syncwarp();
  return lookback_resolve(desc, tile, aggregate);
}

// FLANN L2_Simple<float> restated: fp32, separate multiplies and adds, x then y then z.
__device__ __forceinline__ float l2_simple(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  float r = __fmul_rn(dx, dx);
  r = __fadd_rn(r, __fmul_rn(dy, dy));
  r = __fadd_rn(r, __fmul_rn(dz, dz));
  return r;
}

}  // namespace cp
