// radix_sort.cuh — hand-written stable LSD radix sort of (u64 key, u32 value) pairs.
//
// No CUB / Thrust.  8-bit digits, three kernels per pass (tile histogram, per-digit scan
// over tiles, stable scatter).  The element count and the number of significant key bits
// are read from device memory, so a sort whose size depends on earlier kernels needs no
// host synchronisation: the host enqueues ceil(max_bits/8) passes and passes beyond the
// live bit count exit immediately.  Data ping-pongs A -> B -> A ...; after the sort the
// result sits in sorted_in_b(bits) ? B : A.
#pragma once
#include "common.cuh"

namespace cp {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRounds = 8;                                   // items per thread
constexpr int kSortTile = kSortThreads * kSortRounds;            // 2048 items per tile
constexpr int kRadix = 256;

__host__ __device__ __forceinline__ u32 sort_passes(u32 bits) { return (bits + 7u) >> 3; }
__host__ __device__ __forceinline__ bool sorted_in_b(u32 bits) { return sort_passes(bits) & 1u; }

struct SortArgs {
  u64* keys_a;
  u64* keys_b;
  u32* vals_a;
  u32* vals_b;
  const u32* d_n;     // device: element count
  const u32* d_bits;  // device: significant key bits
  u32* tile_hist;     // [256][tiles_cap] digit-major per-tile counts -> exclusive offsets
  u32* digit_total;   // [256]
  u32 tiles_cap;
};

// pass kernel 1: per-tile digit histogram
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(SortArgs a, u32 pass) {
  const u32 bits = *a.d_bits, n = *a.d_n;
  if (pass * 8 >= bits || n == 0) return;
  const u32 tiles = (n + kSortTile - 1) / kSortTile;
  const u64* src = (pass & 1) ? a.keys_b : a.keys_a;
  const u32 shift = pass * 8;
  __shared__ u32 hist[kRadix];
  for (u32 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    hist[threadIdx.x] = 0;
    __syncthreads();
    const u32 base = tile * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
      const u32 i = base + r * kSortThreads + threadIdx.x;
      if (i < n) atomicAdd(&hist[(u32)(src[i] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    a.tile_hist[(u64)threadIdx.x * tiles + tile] = hist[threadIdx.x];
    __syncthreads();
  }
}

// pass kernel 2: one CTA per digit — exclusive scan of that digit's counts over tiles
__global__ void __launch_bounds__(1024) sort_scan_kernel(SortArgs a, u32 pass) {
  const u32 bits = *a.d_bits, n = *a.d_n;
  if (pass * 8 >= bits || n == 0) return;
  const u32 tiles = (n + kSortTile - 1) / kSortTile;
  u32* row = a.tile_hist + (u64)blockIdx.x * tiles;
  __shared__ u32 warp_sum[32];
  __shared__ u32 carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  for (u32 c0 = 0; c0 < tiles; c0 += blockDim.x) {
    const u32 i = c0 + threadIdx.x;
    const u32 v = (i < tiles) ? row[i] : 0u;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      u32 t = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      u32 w = warp_sum[lane];
      u32 winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(kFull, winc, o);
        if (lane >= o) winc += t;
      }
      warp_sum[lane] = winc - w;  // exclusive over warps
    }
    __syncthreads();
    const u32 carry = carry_s;
    const u32 excl = carry + warp_sum[warp] + inc - v;
    if (i < tiles) row[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) a.digit_total[blockIdx.x] = carry_s;
}

// pass kernel 3: stable scatter.  Warp w of a tile owns the contiguous items
// [w*256, (w+1)*256) and walks them 32 at a time, so ranks follow input order.
__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(SortArgs a, u32 pass) {
  const u32 bits = *a.d_bits, n = *a.d_n;
  if (pass * 8 >= bits || n == 0) return;
  const u32 tiles = (n + kSortTile - 1) / kSortTile;
  const bool odd = pass & 1;
  const u64* ksrc = odd ? a.keys_b : a.keys_a;
  u64* kdst = odd ? a.keys_a : a.keys_b;
  const u32* vsrc = odd ? a.vals_b : a.vals_a;
  u32* vdst = odd ? a.vals_a : a.vals_b;
  const u32 shift = pass * 8;

  __shared__ u32 whist[kSortWarps][kRadix];
  __shared__ u32 gbase[kRadix];
  const int lane = lane_id(), warp = threadIdx.x >> 5;

  // exclusive scan of the 256 digit totals (every CTA recomputes it; 256 values)
  {
    u32 v = a.digit_total[threadIdx.x];
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      u32 t = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += t;
    }
    __shared__ u32 wtot[kSortWarps];
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    u32 off = 0;
    for (int w = 0; w < warp; ++w) off += wtot[w];
    gbase[threadIdx.x] = off + inc - v;
    __syncthreads();
  }

  for (u32 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) whist[w][threadIdx.x] = 0;
    __syncthreads();
    const u32 base = tile * kSortTile + warp * (32 * kSortRounds);
    u64 key[kSortRounds];
    u32 off[kSortRounds];
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
      const u32 i = base + r * 32 + lane;
      const bool valid = i < n;
      key[r] = valid ? ksrc[i] : 0ull;
      const u32 d = valid ? ((u32)(key[r] >> shift) & 0xFFu) : 0x100u;
      const u32 peers = __match_any_sync(kFull, d);
      const int leader = __ffs(peers) - 1;
      u32 old = 0;
      if (valid && lane == leader) {
        old = whist[warp][d];
        whist[warp][d] = old + __popc(peers);
      }
      old = __shfl_sync(kFull, old, leader);
      off[r] = old + __popc(peers & lanemask_lt());
      __syncwarp();
    }
    __syncthreads();
    // exclusive scan across the warps of the tile, per digit
    {
      u32 run = 0;
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        u32 t = whist[w][threadIdx.x];
        whist[w][threadIdx.x] = run;
        run += t;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
      const u32 i = base + r * 32 + lane;
      if (i < n) {
        const u32 d = (u32)(key[r] >> shift) & 0xFFu;
        const u32 pos = gbase[d] + a.tile_hist[(u64)d * tiles + tile] + whist[warp][d] + off[r];
        kdst[pos] = key[r];
        vdst[pos] = vsrc[i];
      }
    }
    __syncthreads();
  }
}

// enqueue every pass the host-side upper bound on the key width can need
inline int radix_sort_enqueue(cudaStream_t st, const SortArgs& a, u32 max_bits, u32 max_n, int sms) {
  const u32 passes = sort_passes(max_bits);
  u32 tiles = (max_n + kSortTile - 1) / kSortTile;
  if (tiles == 0) tiles = 1;
  u32 grid = tiles < (u32)(sms * 8) ? tiles : (u32)(sms * 8);
  int launches = 0;
  for (u32 p = 0; p < passes; ++p) {
    sort_hist_kernel<<<grid, kSortThreads, 0, st>>>(a, p);
    sort_scan_kernel<<<kRadix, 1024, 0, st>>>(a, p);
    sort_scatter_kernel<<<grid, kSortThreads, 0, st>>>(a, p);
    launches += 3;
  }
  return launches;
}

}  // namespace cp
