// radix_sort.cuh — hand-written stable LSD radix sort of (u64 key, u32 value) pairs.
//
// No CUB / Thrust.  8-bit digits, ONE kernel per pass ("onesweep"): a tile ranks its 2048 items per digit
// (stable: warp-contiguous chunks + match_any), publishes its 256 digit counts, finds the counts of the tiles
// before it by decoupled look-back (tiles are handed out by ticket, so every earlier tile is already running) and
// leaves through shared memory in sorted order, so runs of equal digits store coalesced.  The global digit bases
// of ALL passes are counted once — the key multiset does not change between passes — either by the kernel that
// writes the keys (sort_feed_*: the general back half's four sorts, which are then their passes and nothing else)
// or by sort_prepare_kernel (a sort of keys that are already there: 1 small memset + 1 + passes launches).
// The element count and the number of significant key bits are read from device memory, so a sort whose size
// depends on earlier kernels needs no host synchronisation: the host enqueues ceil(max_bits/8) passes and passes
// beyond the live bit count exit immediately.  Data ping-pongs A -> B -> A ...; after the sort the
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

constexpr u32 kSortMaxPasses = 8;
constexpr u32 kSortHdrWords = kSortMaxPasses + kSortMaxPasses * kRadix;   // tickets, then digit totals per pass
constexpr u32 kFlagAgg = 1u << 30, kFlagPre = 2u << 30, kSortValMask = (1u << 30) - 1u;

struct SortArgs {
  u64* keys_a;
  u64* keys_b;
  u32* vals_a;
  u32* vals_b;
  const u32* d_n;     // device: element count
  const u32* d_bits;  // device: significant key bits
  u32* hdr;           // [8] tile tickets per pass, [8][256] digit totals per pass (zeroed by the host per sort)
  u32* state;         // [2][tiles_cap][256] look-back words: flag | count, double-buffered by pass parity
  u32 tiles_cap;
};

// once per sort: digit histograms of every live pass (one read of the keys) and a clean look-back buffer for pass 0
__global__ void __launch_bounds__(kSortThreads) sort_prepare_kernel(SortArgs a) {
  const u32 bits = *a.d_bits, n = *a.d_n;
  if (bits == 0 || n == 0) return;
  const u32 passes = sort_passes(bits) < kSortMaxPasses ? sort_passes(bits) : kSortMaxPasses;
  const u32 tiles = (n + kSortTile - 1) / kSortTile;
  __shared__ u32 hist[kSortMaxPasses][kRadix];
  for (u32 p = 0; p < passes; ++p) hist[p][threadIdx.x] = 0;
  __syncthreads();
  for (u32 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    a.state[(u64)tile * kRadix + threadIdx.x] = 0u;
    const u32 base = tile * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
      const u32 i = base + r * kSortThreads + threadIdx.x;
      if (i < n) {
        const u64 k = a.keys_a[i];
        for (u32 p = 0; p < passes; ++p) atomicAdd(&hist[p][(u32)(k >> (8 * p)) & 0xFFu], 1u);
      }
    }
  }
  __syncthreads();
  for (u32 p = 0; p < passes; ++p) {
    const u32 c = hist[p][threadIdx.x];
    if (c) atomicAdd(&a.hdr[kSortMaxPasses + p * kRadix + threadIdx.x], c);
  }
}

// ---- producer side of a sort whose keys come from one of this library's kernels -------------------------------
// The kernel that writes keys_a also counts their digits for every live pass and clears the look-back words of
// pass 0 — the work of sort_prepare_kernel without a launch (and a second read of the keys) of its own.  The host
// clears the header before the producer (radix_sort_clear_hdr) and enqueues the passes behind it
// (radix_sort_passes).  Every thread of the CTA must call begin / flush; sort_feed_key is called by whole warps
// (`valid` = this lane has a key).
struct SortFeedSmem {
  u32 hist[kSortMaxPasses][kRadix];
};
__device__ __forceinline__ u32 sort_feed_passes(u32 bits, u32 n) {
  if (bits == 0 || n == 0) return 0;
  return sort_passes(bits) < kSortMaxPasses ? sort_passes(bits) : kSortMaxPasses;
}
__device__ __forceinline__ void sort_feed_begin(SortFeedSmem& s, u32 passes) {
  for (u32 i = threadIdx.x; i < passes * kRadix; i += blockDim.x) (&s.hist[0][0])[i] = 0;
  __syncthreads();
}
__device__ __forceinline__ void sort_feed_key(SortFeedSmem& s, u64 key, bool valid, u32 passes) {
  for (u32 p = 0; p < passes; ++p) {
    // neighbouring keys share their high digits: one shared-memory atomic per distinct digit of the warp
    const u32 d = valid ? ((u32)(key >> (8 * p)) & 0xFFu) : 0x100u;
    const u32 peers = __match_any_sync(kFull, d);
    if (valid && lane_id() == __ffs(peers) - 1) atomicAdd(&s.hist[p][d], (u32)__popc(peers));
  }
}
__device__ __forceinline__ void sort_feed_flush(SortFeedSmem& s, u32* hdr, u32* state, u32 passes, u32 n) {
  __syncthreads();
  for (u32 i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
    const u32 c = (&s.hist[0][0])[i];
    if (c) atomicAdd(&hdr[kSortMaxPasses + i], c);
  }
  if (passes) {
    const u32 words = ((n + kSortTile - 1) / kSortTile) * kRadix;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) state[i] = 0u;
  }
}

// one pass.  Warp w of a tile owns the contiguous items [w*256, (w+1)*256) and walks them 32 at a time, so ranks
// follow input order (stable).
__global__ void __launch_bounds__(kSortThreads, 4) sort_onesweep_kernel(SortArgs a, u32 pass) {
  __shared__ u32 whist[kSortWarps][kRadix];
  __shared__ u32 gbase[kRadix];
  __shared__ u32 lbase[kRadix];          // first slot of a digit inside the tile's staged (sorted) order
  __shared__ u32 dbase[kRadix];          // global position of staged slot i of digit d = dbase[d] + i
  __shared__ u32 wsum2[kSortWarps];
  __shared__ u64 st_key[kSortTile];      // the tile in sorted order: runs of equal digits leave with coalesced
  __shared__ u32 st_val[kSortTile];      // stores (scattered from registers, every 8-byte store was its own sector)
  __shared__ u32 s_tile;
  // the first ticket, the sizes and this pass's digit totals are independent loads: all in flight together (one
  // L2 round trip instead of three in front of a tile's first key load — a small sort is one wave of tiles, and
  // its passes are as long as that chain)
  if (threadIdx.x == 0) s_tile = atomicAdd(&a.hdr[pass], 1u);   // ticket: earlier tiles are already running
  u32 v = a.hdr[kSortMaxPasses + pass * kRadix + threadIdx.x];
  const u32 bits = *a.d_bits, n = *a.d_n;
  if (pass * 8 >= bits || n == 0) return;
  const u32 tiles = (n + kSortTile - 1) / kSortTile;
  const bool odd = pass & 1;
  const u64* ksrc = odd ? a.keys_b : a.keys_a;
  u64* kdst = odd ? a.keys_a : a.keys_b;
  const u32* vsrc = odd ? a.vals_b : a.vals_a;
  u32* vdst = odd ? a.vals_a : a.vals_b;
  const u32 shift = pass * 8;
  u32* st = a.state + (u64)(pass & 1u) * a.tiles_cap * kRadix;
  u32* st_next = a.state + (u64)((pass + 1u) & 1u) * a.tiles_cap * kRadix;
  const int lane = lane_id(), warp = threadIdx.x >> 5;

  // exclusive scan of this pass's 256 digit totals (every CTA recomputes it; 256 values)
  {
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
  }

  bool first = true;
  while (true) {
    __syncthreads();
    if (!first && threadIdx.x == 0) s_tile = atomicAdd(&a.hdr[pass], 1u);
    first = false;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) whist[w][threadIdx.x] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    if (tile >= tiles) break;
    const u32 base = tile * kSortTile + warp * (32 * kSortRounds);
    u64 key[kSortRounds];
    u32 val[kSortRounds];
    u32 off[kSortRounds];
    // all loads first (the values too: fetched before the scatter they would be one more round trip at the end
    // of every tile): the __syncwarp() of the ranking rounds would otherwise put one memory round trip between
    // every two rounds
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
      const u32 i = base + r * 32 + lane;
      key[r] = i < n ? ksrc[i] : 0ull;
      val[r] = i < n ? vsrc[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
      const u32 i = base + r * 32 + lane;
      const bool valid = i < n;
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
    // per digit (thread = digit): exclusive scan across the warps of the tile, then the tiles before this one
    {
      u32 run = 0;
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        u32 t = whist[w][threadIdx.x];
        whist[w][threadIdx.x] = run;
        run += t;
      }
      // where the digit's run starts inside the tile: exclusive scan of the 256 runs (finished after the look-back)
      u32 inc2 = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(kFull, inc2, o);
        if (lane >= o) inc2 += t;
      }
      if (lane == 31) wsum2[warp] = inc2;
      volatile u32* mine = st + (u64)tile * kRadix + threadIdx.x;
      u32 ex = 0;
      if (tile == 0) {
        *mine = kFlagPre | run;
      } else {
        *mine = kFlagAgg | run;
        // look-back, kLook predecessors per round trip (independent loads in flight together): counts of tiles
        // that have only published their own aggregate are added and the walk goes on; the first inclusive
        // prefix ends it (tile 0 always publishes one).  A word that is not written yet is read again.
        const volatile u32* col = st + threadIdx.x;
        i32 t = (i32)tile - 1;
        bool done = false;
        while (!done) {
          constexpr int kLook = 8;
          u32 sv[kLook];
#pragma unroll
          for (int q = 0; q < kLook; ++q) sv[q] = (t - q >= 0) ? col[(u64)(t - q) * kRadix] : kFlagPre;
          int consumed = 0;
#pragma unroll
          for (int q = 0; q < kLook; ++q) {
            if (done || consumed != q) continue;      // stop at the first word that is not ready
            if (sv[q] & kFlagPre) {
              ex += sv[q] & kSortValMask;
              done = true;
            } else if (sv[q] & kFlagAgg) {
              ex += sv[q] & kSortValMask;
              ++consumed;
            }
          }
          t -= consumed;
        }
        *mine = kFlagPre | (ex + run);
      }
      st_next[(u64)tile * kRadix + threadIdx.x] = 0u;   // clean look-back words for the next pass
      __syncthreads();
      u32 tstart = inc2 - run;
      for (int w = 0; w < warp; ++w) tstart += wsum2[w];
      lbase[threadIdx.x] = tstart;
      dbase[threadIdx.x] = gbase[threadIdx.x] + ex - tstart;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
      const u32 i = base + r * 32 + lane;
      if (i < n) {
        const u32 d = (u32)(key[r] >> shift) & 0xFFu;
        const u32 slot = lbase[d] + whist[warp][d] + off[r];
        st_key[slot] = key[r];
        st_val[slot] = val[r];
      }
    }
    __syncthreads();
    const u32 in_tile = n - tile * kSortTile < (u32)kSortTile ? n - tile * kSortTile : (u32)kSortTile;
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
      const u32 i = r * kSortThreads + threadIdx.x;
      if (i < in_tile) {
        const u64 kk = st_key[i];
        const u32 pos = dbase[(u32)(kk >> shift) & 0xFFu] + i;
        kdst[pos] = kk;
        vdst[pos] = st_val[i];
      }
    }
  }
}

// the passes alone, behind a producer kernel that fed the histograms (sort_feed_*)
inline int radix_sort_passes(cudaStream_t st, const SortArgs& a, u32 max_bits, u32 max_n, int sms, bool persistent) {
  u32 passes = sort_passes(max_bits);
  if (passes > kSortMaxPasses) passes = kSortMaxPasses;
  u32 tiles = (max_n + kSortTile - 1) / kSortTile;
  if (tiles == 0) tiles = 1;
  // (one CTA per tile up to 16 per SM: a handle's first run has no size hint yet, and launching a CTA for every
  // tile of the CAPACITY cost 12 us per pass on a sort of a few thousand items; CTAs loop over tickets anyway)
  const u32 cap = (u32)sms * (persistent ? 3u : 16u);
  const u32 grid = tiles < cap ? tiles : cap;
  for (u32 p = 0; p < passes; ++p) sort_onesweep_kernel<<<grid, kSortThreads, 0, st>>>(a, p);
  return (int)passes;
}

// enqueue every pass the host-side upper bound on the key width can need
inline int radix_sort_enqueue(cudaStream_t st, const SortArgs& a, u32 max_bits, u32 max_n, int sms,
                              bool persistent = false) {
  u32 passes = sort_passes(max_bits);
  if (passes > kSortMaxPasses) passes = kSortMaxPasses;
  u32 tiles = (max_n + kSortTile - 1) / kSortTile;
  if (tiles == 0) tiles = 1;
  // tiles by ticket; one CTA per tile unless `persistent` (see cp_handle::tile_ctas in pipeline.cu): a CTA that
  // goes on to another tile cannot publish that tile's counts before its current look-back has resolved
  const u32 gcap = (u32)sms * (persistent ? 3u : 16u);
  u32 grid = tiles < gcap ? tiles : gcap;
  cudaMemsetAsync(a.hdr, 0, sizeof(u32) * kSortHdrWords, st);
  const u32 pgrid = tiles < (u32)(sms * 3) ? tiles : (u32)(sms * 3);   // histograms: grid-stride, no look-back
  sort_prepare_kernel<<<pgrid, kSortThreads, 0, st>>>(a);
  int launches = 1;
  for (u32 p = 0; p < passes; ++p) {
    sort_onesweep_kernel<<<grid, kSortThreads, 0, st>>>(a, p);
    ++launches;
  }
  return launches;
}

}  // namespace cp
