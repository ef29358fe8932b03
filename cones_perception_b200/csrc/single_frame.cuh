// single_frame.cuh — a node's duty cycle in ONE launch: one PointCloud2 in, the cone list out.
//
// The reference nodes handle one frame per callback (ros::spin, subscriber queue 2: src/cone_detection.cpp:112,127;
// src/ground_removal.cpp:41,47).  For a single frame the multi-launch pipeline (reset, pass 1, per-frame kernel,
// result publish) is dominated by launch gaps and by one CTA doing the second pass over the whole scan on its own.
// Here a thread-block cluster of 16 CTAs owns the frame:
//   1. every CTA streams its 1/16 of the frame from HBM once, updating its sector minima on the fly (pass 1 of
//      src/ground_removal.cpp:58-68) and stashing x, y, z in its shared memory;
//   2. cluster barrier; the 16 partial sector tables are combined through distributed shared memory;
//   3. pass 2 (ground verdicts :70-77 + the crop of src/cone_detection.cpp:189-204) runs out of the stash and
//      writes the keep-mask words and per-tile survivor counts;
//   4. cluster barrier; 15 CTAs leave, CTA 0 runs the whole back half (frame_process: VoxelGrid, clustering,
//      centroids — frame_kernels.cuh) with its shared memory laid over the dead stash, and
//   5. stores control block, counters, offsets and the K cone records straight into the pinned host mirrors.
// Nothing has to be reset beforehand (every word the kernel reads it has written itself), so a frame is exactly
// one kernel launch and one stream synchronisation.  Without ground removal steps 1-2 disappear (the points are
// judged as they arrive).  Results are the multi-launch path's, bit for bit (same device functions).
#pragma once
#include <cooperative_groups.h>

#include "frame_kernels.cuh"

namespace cp {

constexpr int kSfCluster = 16;
constexpr int kSfThreads = 512;
constexpr int kSfWarps = kSfThreads / 32;
constexpr u32 kSfSub = kSfThreads * kStreamRows;      // 4096 points per CTA round = two mask tiles
constexpr u32 kSfMaxPtsPerCta = 16384;                // 192 KB stash

struct SingleArgs {
  FrameArgs fa;        // the back half of frame 0 (mask = keep-mask words written by step 3)
  u32 n;               // points of the frame
  u32 pts_per_cta;     // multiple of kSfSub
  float default_low;
  MaskOut o;           // mask words, per-tile survivor counts, ground survivors
  u32* low_key;        // [32] final sector minima (counters / cp_ground_remove's low17)
  // result publish: pinned host mirrors (device-visible)
  u32* h_ctl;
  u32* h_fc;
  u32* h_res;
  u32 off_words, max_records;
};

template <int CMAX, int VMAX, int MODE>
__global__ void __launch_bounds__(kSfThreads, 1) single_frame_kernel(const __grid_constant__ SingleArgs a) {
  namespace cg = cooperative_groups;
  extern __shared__ __align__(16) unsigned char sf_smem[];
  float* sx = reinterpret_cast<float*>(sf_smem);
  float* sy = sx + a.pts_per_cta;
  float* sz = sy + a.pts_per_cta;
  __shared__ u32 smin[kSectStride];      // this CTA's sector table (read remotely after the first barrier)
  __shared__ u32 red[kSectStride];
  __shared__ u32 s_bound[5];
  __shared__ float thr[kSectStride];     // [0..16] thresholds, [31] their minimum, [30] their maximum
  __shared__ u32 wtot[kSfWarps];
  __shared__ u32 s_gkept;                // ground survivors: every CTA adds its part into CTA 0's copy

  cg::cluster_group cluster = cg::this_cluster();
  const u32 crank = cluster.block_rank();
  const u32 tid = threadIdx.x;
  const int lane = lane_id(), warp = tid >> 5;
  const u32 p0 = crank * a.pts_per_cta;
  const u32 cnt = p0 < a.n ? min(a.n - p0, a.pts_per_cta) : 0u;
  const u32 nsub = a.pts_per_cta / kSfSub;
  const u32 wl = (u32)warp * (32 * kStreamRows) + lane;   // this thread's first point in a round
  const bool do_ground = a.fa.gk.do_ground != 0;
  const float kInf = __int_as_float(0x7f800000);

  if (tid < kSectStride) {
    smin[tid] = tid < kNSect ? f2ord(a.default_low) : 0u;
    red[tid] = 0xFFFFFFFFu;
  }
  if (tid == 0) s_gkept = 0;
  if (!do_ground && tid < kSectStride) thr[tid] = -kInf;
  __syncthreads();

  u32 gkept = 0;
  if (do_ground) {
    // ---- 1: stream once (next round's loads in flight), minima on the fly, x/y/z into the stash
    float4 cur[kStreamRows], nxt[kStreamRows];
#pragma unroll
    for (int r = 0; r < kStreamRows; ++r) {
      const u32 i = wl + r * 32;
      cur[r] = (i < cnt) ? load_point<MODE>(a.fa.in, (u64)p0 + i, a.fa.layout) : make_float4(0.f, 0.f, kInf, 0.f);
    }
    for (u32 sub = 0; sub < nsub; ++sub) {
      if (sub + 1 < nsub) {
#pragma unroll
        for (int r = 0; r < kStreamRows; ++r) {
          const u32 i = (sub + 1) * kSfSub + wl + r * 32;
          nxt[r] = (i < cnt) ? load_point<MODE>(a.fa.in, (u64)p0 + i, a.fa.layout) : make_float4(0.f, 0.f, kInf, 0.f);
        }
      }
      sector_bounds(smin, s_bound);   // bounds from the minima seen so far (barrier inside)
#pragma unroll
      for (int r = 0; r < kStreamRows; ++r) {
        const u32 i = sub * kSfSub + wl + r * 32;
        sx[i] = cur[r].x;
        sy[i] = cur[r].y;
        sz[i] = cur[r].z;
      }
      sector_min_tile(cur, smin, s_bound);
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kStreamRows; ++r) cur[r] = nxt[r];
    }
    // ---- 2: combine the 16 partial tables through distributed shared memory
    cluster.sync();
    for (u32 t = tid; t < kNSect * kSfCluster; t += kSfThreads) {
      const u32 s = t % kNSect, rr = t / kNSect;
      atomicMin(&red[s], *cluster.map_shared_rank(&smin[s], rr));
    }
    __syncthreads();
    if (warp == 0) {
      // :75  p.z < low + 0.1 in double  <=>  z < roundup_to_float((double)low + 0.1)
      float t = 0.f;
      if (lane < kNSect) {
        const u32 m = red[lane];
        if (crank == 0) a.low_key[lane] = m;
        t = __double2float_ru((double)ord2f(m) + 0.1);
        thr[lane] = t;
      }
      float mn = lane < kNSect ? t : kInf, mx = lane < kNSect ? t : -kInf;
#pragma unroll
      for (int o2 = 16; o2; o2 >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(kFull, mn, o2));
        mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o2));
      }
      if (lane == 0) {
        thr[31] = mn;
        thr[30] = mx;
      }
    }
    __syncthreads();
  }
  const float thr_min = thr[31], thr_max = thr[30];

  // ---- 3: keep bits + survivors per mask tile (out of the stash, or straight from HBM without ground removal)
  for (u32 sub = 0; sub < nsub; ++sub) {
    u32 wcount = 0, myword = 0;
    float4 q[kStreamRows];
#pragma unroll
    for (int r = 0; r < kStreamRows; ++r) {
      const u32 i = sub * kSfSub + wl + r * 32;
      if (do_ground) q[r] = make_float4(sx[i], sy[i], sz[i], 0.f);
      else q[r] = (i < cnt) ? load_point<MODE>(a.fa.in, (u64)p0 + i, a.fa.layout) : make_float4(0.f, 0.f, -kInf, 0.f);
    }
#pragma unroll
    for (int r = 0; r < kStreamRows; ++r) {
      const u32 i = sub * kSfSub + wl + r * 32;
      const u32 bal = __ballot_sync(kFull, keep_point(q[r], i < cnt, a.fa.crop, a.fa.gk, thr, thr_min, thr_max, gkept));
      wcount += __popc(bal);
      if (lane == r) myword = bal;
    }
    // mask words are linear in the frame: word w covers points 32 w .. 32 w + 31
    const u32 first_word = (p0 + sub * kSfSub) / 32 + (u32)warp * kStreamRows;
    const u32 n_words = (a.n + 31u) / 32u;
    if (lane < kStreamRows && first_word + lane < a.fa.geom.tpf * kTileWords)
      a.o.mask[first_word + lane] = first_word + lane < n_words ? myword : 0u;
    if (lane == 0) wtot[warp] = wcount;
    __syncthreads();
    if ((tid & 255u) == 0) {   // warps 0-7 fill one 2048-point mask tile, warps 8-15 the next
      const u32 half = tid >> 8;
      const u32 tile = (p0 + sub * kSfSub) / kStreamTile + half;
      if (tile < a.fa.geom.tpf) {
        u32 total = 0;
#pragma unroll
        for (int w = 0; w < kStreamWarps; ++w) total += wtot[half * kStreamWarps + w];
        a.o.tile_count[tile] = total + ((a.fa.gk.pad_survives && tile + 1 == a.fa.geom.tpf) ? 1u : 0u);
      }
    }
    __syncthreads();
  }
  if (do_ground) {
    gkept = __reduce_add_sync(kFull, gkept);
    if (lane == 0 && gkept) atomicAdd(cluster.map_shared_rank(&s_gkept, 0), gkept);
  }

  // ---- 4: everything the back half reads is written; 15 CTAs are done
  cluster.sync();
  if (crank != 0) return;
  if (tid == 0) {
    Ctl z;
    memset(&z, 0, sizeof(z));
    *a.fa.ctl = z;
    a.o.gcount[0] = s_gkept;
  }
  __syncthreads();
  FrameSmem<CMAX, VMAX, kSfThreads>& s = *reinterpret_cast<FrameSmem<CMAX, VMAX, kSfThreads>*>(sf_smem);
  frame_process<CMAX, VMAX, MODE, kSfThreads>(a.fa, s, 0u);

  // ---- 5: results into the pinned host mirrors (zero-copy stores; what result_publish_kernel does)
  __threadfence();
  __syncthreads();
  const u32 K = min(a.fa.ctl->n_clusters, a.max_records);
  for (u32 i = tid; i < sizeof(Ctl) / 4; i += kSfThreads) a.h_ctl[i] = reinterpret_cast<const u32*>(a.fa.ctl)[i];
  if (tid < 8) a.h_fc[tid] = a.fa.fc[tid];
  const u32 res_words = a.off_words + 4u * K;
  for (u32 i = tid; i < res_words; i += kSfThreads) a.h_res[i] = a.fa.k_off[i];
}

}  // namespace cp
