// cluster_front.cuh — the whole streaming front end in ONE pass over HBM.
//
// A thread-block cluster of 16 CTAs owns one frame at a time.  Each CTA streams its 1/16 of the
// frame from HBM exactly once, updating its sector minima on the fly (pass 1 of
// src/ground_removal.cpp:58-68) and stashing x, y, z (12 B/point, SoA) in its shared memory.
// The 16 partial sector tables are combined through distributed shared memory after one cluster
// barrier; pass 2 (src/ground_removal.cpp:70-77 fused with src/cone_detection.cpp:189-204) then
// runs entirely out of shared memory and writes the same keep-mask words / tile counts as
// keep_mask_kernel.  HBM traffic: 16 B per point, once — half of the two-kernel front end.
// With frames of <= 131 072 points the stash is 96 KB, so two CTAs (two clusters) share an SM
// and one cluster's shared-memory phases overlap the other's HBM streaming.
//
// Used for uniform batches in the compact float4 layout with ground removal on, frames up to
// 16 x 16384 points; everything else takes the two-kernel path (same results).
#pragma once
#include <cooperative_groups.h>

#include "stream_kernels.cuh"

namespace cp {

namespace cg = cooperative_groups;

constexpr int kClusterSize = 16;
constexpr int kClThreads = kStreamThreads;           // 256: one sub-tile = one mask tile of 2048 points
constexpr u32 kClMaxPtsPerCta = 16384;               // 192 KB stash (one CTA per SM at that size)

struct ClusterArgs {
  const float4* in;
  u32 n_frames;
  u32 n;             // points per frame (uniform)
  u32 pts_per_cta;   // multiple of kStreamTile
  u32 tpf;           // mask tiles (2048 points) per frame
  float default_low;
  CropK c;
  GroundK gk;
  u32* low_key;      // [F][32] final minima (for counters / taps)
  MaskOut o;
};

__global__ void __launch_bounds__(kClThreads, 2) front_cluster_kernel(ClusterArgs a) {
  extern __shared__ __align__(16) unsigned char cl_smem[];
  float* sx = reinterpret_cast<float*>(cl_smem);
  float* sy = sx + a.pts_per_cta;
  float* sz = sy + a.pts_per_cta;
  __shared__ u32 smin[2][kSectStride];   // per-CTA sector table, double-buffered by frame parity (read remotely)
  __shared__ u32 red[kSectStride];
  __shared__ u32 s_bound[5];
  __shared__ float thr[kSectStride];
  __shared__ float s_thr_min;
  __shared__ u32 wtot[kStreamWarps];

  cg::cluster_group cluster = cg::this_cluster();
  const u32 crank = cluster.block_rank();
  const u32 n_clusters = gridDim.x / kClusterSize;
  const u32 cid = blockIdx.x / kClusterSize;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const u32 p0 = crank * a.pts_per_cta;                                  // first point of this CTA in the frame
  const u32 cnt = p0 < a.n ? min(a.n - p0, a.pts_per_cta) : 0u;          // points this CTA owns
  const u32 nsub = a.pts_per_cta / kStreamTile;
  const u32 dkey = f2ord(a.default_low);
  const u32 wl = warp * (32 * kStreamRows) + lane;                       // this thread's first point in a sub-tile
  u32 parity = 0;

  for (u32 frame = cid; frame < a.n_frames; frame += n_clusters, parity ^= 1u) {
    u32* tbl = smin[parity];
    if (threadIdx.x < kSectStride) {
      tbl[threadIdx.x] = threadIdx.x < kNSect ? dkey : 0u;
      red[threadIdx.x] = 0xFFFFFFFFu;
    }
    const float4* src = a.in + (u64)frame * a.n + p0;

    // ---- pass 1: stream from HBM once (next sub-tile's loads in flight while this one is
    // processed), minima on the fly, x/y/z into the stash
    float4 cur[kStreamRows], nxt[kStreamRows];
#pragma unroll
    for (int r = 0; r < kStreamRows; ++r) {
      const u32 i = wl + r * 32;
      cur[r] = (i < cnt) ? ldg_stream(src + i) : make_float4(0.f, 0.f, __int_as_float(0x7f800000), 0.f);
    }
    __syncthreads();
    for (u32 sub = 0; sub < nsub; ++sub) {
      if (sub + 1 < nsub) {
#pragma unroll
        for (int r = 0; r < kStreamRows; ++r) {
          const u32 i = (sub + 1) * kStreamTile + wl + r * 32;
          nxt[r] = (i < cnt) ? ldg_stream(src + i) : make_float4(0.f, 0.f, __int_as_float(0x7f800000), 0.f);
        }
      }
      sector_bounds(tbl, s_bound);   // bounds from the minima seen so far (barrier inside)
#pragma unroll
      for (int r = 0; r < kStreamRows; ++r) {
        const u32 i = sub * kStreamTile + wl + r * 32;
        sx[i] = cur[r].x;
        sy[i] = cur[r].y;
        sz[i] = cur[r].z;
      }
      sector_min_tile(cur, tbl, s_bound);
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kStreamRows; ++r) cur[r] = nxt[r];
    }

    // ---- combine the 16 partial tables through distributed shared memory
    cluster.sync();
    for (u32 t = threadIdx.x; t < kNSect * kClusterSize; t += kClThreads) {
      const u32 s = t % kNSect, rr = t / kNSect;
      atomicMin(&red[s], *cluster.map_shared_rank(&smin[parity][s], rr));
    }
    __syncthreads();
    if (threadIdx.x < kNSect) {
      const u32 m = red[threadIdx.x];
      if (crank == 0) a.low_key[frame * kSectStride + threadIdx.x] = m;
      // :75  p.z < low + 0.1 in double  <=>  z < roundup_to_float((double)low + 0.1)
      thr[threadIdx.x] = __double2float_ru((double)ord2f(m) + 0.1);
    }
    __syncthreads();
    if (warp == 0) {
      float t = lane < kNSect ? thr[lane] : __int_as_float(0x7f800000);
#pragma unroll
      for (int o2 = 16; o2; o2 >>= 1) t = fminf(t, __shfl_xor_sync(kFull, t, o2));
      if (lane == 0) s_thr_min = t;
    }
    __syncthreads();
    const float thr_min = s_thr_min;

    // ---- pass 2 out of shared memory: keep bits + survivors per mask tile
    for (u32 sub = 0; sub < nsub; ++sub) {
      u32 wcount = 0, gkept = 0, myword = 0;
#pragma unroll
      for (int r = 0; r < kStreamRows; ++r) {
        const u32 i = sub * kStreamTile + wl + r * 32;
        const float z = sz[i];
        bool keep = false;
        if ((i < cnt) && !(z < thr_min)) {
          const float x = sx[i], y = sy[i];
          if (finite3(x, y, z)) {
            keep = !a.c.do_crop || crop_keep(a.c, x, y, z);
            if (keep || a.gk.want_count) {
              bool ok;
              const float ang = atan2_approx(y, x, ok);
              if (keep && a.c.do_crop) {
                const float aa = fabsf(ang);
                if (!ok | (aa > a.c.f_lo_guard)) {
                  if (ok & (aa >= a.c.f_hi_guard)) keep = false;
                  else keep = fabsf(atan2_exact(y, x)) < a.c.f_hi;  // src/cone_detection.cpp:200-201
                }
              }
              const int s = sector_of(x, y, ang, ok);
              const bool gkeep = !(z < thr[s]);
              if (gkeep) gkept++;
              keep = keep && gkeep;
            }
          }
        }
        const u32 bal = __ballot_sync(kFull, keep);
        wcount += __popc(bal);
        if (lane == r) myword = bal;
      }
      // one sub-tile = one mask tile; mask words are linear in the frame
      const u32 tile_in_frame = p0 / kStreamTile + sub;
      const bool live = tile_in_frame < a.tpf;
      if (live && lane < kStreamRows)
        a.o.mask[((u64)frame * a.tpf + tile_in_frame) * kTileWords + warp * kStreamRows + lane] = myword;
      if (lane == 0) wtot[warp] = wcount;
      if (a.gk.want_count) {
        const u32 gsum = __reduce_add_sync(kFull, gkept);
        if (lane == 0 && gsum) atomicAdd(&a.o.gcount[frame], gsum);
      }
      __syncthreads();
      if (threadIdx.x == 0 && live) {
        u32 total = 0;
#pragma unroll
        for (int w = 0; w < kStreamWarps; ++w) total += wtot[w];
        a.o.tile_count[frame * a.tpf + tile_in_frame] =
            total + ((a.gk.pad_survives && tile_in_frame + 1 == a.tpf) ? 1u : 0u);
      }
      __syncthreads();
    }
  }
  // nobody may leave while a neighbour can still read its table
  cluster.sync();
}

}  // namespace cp
