// frame_kernels.cuh — the whole back half of the path for one frame per CTA, in shared memory.
//
// After the streaming front end only ~10^2..10^3 points of a 130k-point scan survive, so
// VoxelGrid (src/cone_detection.cpp:240-249), Euclidean clustering with the size filter
// (:206-220) and the centroid loop (:261-273) of one frame fit in one CTA's shared memory:
// no global intermediates, no per-stage launches.  Frames are handed out by ticket and a
// decoupled look-back over frame descriptors packs the per-frame results (voxel and
// cluster offsets) without a second pass.
//
// Semantics are identical to the general (global-memory) path in voxel_kernels.cuh /
// cluster_kernels.cuh; frames that exceed the shared-memory capacity (or trip VoxelGrid's
// int32 overflow guard) raise Ctl::fast_overflow and the host re-runs the back half through
// the general path.
#pragma once
#include "cluster_kernels.cuh"
#include "common.cuh"
#include "voxel_kernels.cuh"

namespace cp {

constexpr int kFrameThreads = 512;

struct FrameArgs {
  u32 n_frames;
  const u32* c_off;          // [F+1]
  const float4* pts;         // survivors
  const u32* src;
  const u32* bbox_key;
  const u32* frame_n;
  u32 uniform_n;
  const u32* gcount;
  int pad_survives;
  VoxelK vk;
  ClusterK ck;
  // outputs
  VoxelFrame* vf;            // [F]
  u32* v_off;                // [F+1]
  u32* k_off;                // [F+1]
  u32* ncomp_f;              // [F]
  u32* kcount_f;             // [F]
  ClusterRec* clusters;      // packed, canonical order per frame
  u32 clusters_cap;
  u64* desc_v;               // frame descriptors for the voxel offsets
  u64* desc_k;               // frame descriptors for the cluster offsets
  Ctl* ctl;
  u32* ticket;               // frame ticket (zeroed before the launch)
  // taps (NULL when off)
  float4* tap_vox;
  u32 tap_vox_cap;
  u32* tap_keys;
  u32* tap_order;
  i32* tap_labels;
};

template <int CMAX, int VMAX>
struct FrameSmem {
  float px[CMAX], py[CMAX], pz[CMAX], pw[CMAX];
  u64 key[CMAX];             // (voxel idx << 32) | survivor position
  float vx[VMAX], vy[VMAX], vz[VMAX];
  u32 vstart[VMAX + 1];      // voxel -> first sorted record; later: kept roots / ranks
  u32 parent[VMAX];
  u32 label[VMAX];
  u32 csize[VMAX];
  u32 wsum[kFrameThreads / 32];
  VoxelFrame vfr;
  u32 frame, slow, n_vox, v_excl, k_excl, n_kept, n_comp;
};

__device__ __forceinline__ u32 smem_find(volatile u32* parent, u32 x) {
  u32 p = parent[x];
  while (p != x) {
    const u32 gp = parent[p];
    if (gp != p) parent[x] = gp;
    x = p;
    p = gp;
  }
  return x;
}
__device__ __forceinline__ void smem_union(u32* parent, u32 a, u32 b) {
  while (true) {
    a = smem_find(parent, a);
    b = smem_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const u32 t = a;
      a = b;
      b = t;
    }
    if (atomicCAS(&parent[a], a, b) == a) return;
  }
}

// block-wide exclusive scan of one value per thread; returns the exclusive prefix and the total
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32* wsum, u32& total) {
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  u32 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // wsum may still be read from a previous call
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  u32 woff = 0;
  total = 0;
#pragma unroll
  for (int w = 0; w < kFrameThreads / 32; ++w) {
    const u32 t = wsum[w];
    if (w < warp) woff += t;
    total += t;
  }
  return woff + inc - v;
}

template <int CMAX, int VMAX>
__global__ void __launch_bounds__(kFrameThreads) frame_backend_kernel(FrameArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FrameSmem<CMAX, VMAX>& s = *reinterpret_cast<FrameSmem<CMAX, VMAX>*>(smem_raw);
  const u32 tid = threadIdx.x;
  const int lane = lane_id(), warp = tid >> 5;

  while (true) {
    __syncthreads();
    if (tid == 0) s.frame = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const u32 f = s.frame;
    if (f >= a.n_frames) break;
    const u32 c0 = a.c_off[f];
    const u32 C = a.c_off[f + 1] - c0;

    // ---- S0: VoxelGrid setup from the survivors' bounding box (thread 0)
    if (tid == 0) {
      VoxelFrame v;
      v.pad = 0;
      v.passthrough = 0;
      v.mul1 = v.mul2 = 0;
      v.bits = 0;
      v.min_b[0] = v.min_b[1] = v.min_b[2] = 0;
      u32 slow = (C > (u32)CMAX) ? 1u : 0u;
      if (C > 0) {
        long long d[3];
        i32 div_b[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float mn = ord2f(a.bbox_key[f * 8 + k]), mx = ord2f(a.bbox_key[f * 8 + 4 + k]);
          d[k] = (long long)__fmul_rn(__fsub_rn(mx, mn), a.vk.inv[k]) + 1;
          v.min_b[k] = (i32)floorf(__fmul_rn(mn, a.vk.inv[k]));
          const i32 max_b = (i32)floorf(__fmul_rn(mx, a.vk.inv[k]));
          div_b[k] = max_b - v.min_b[k] + 1;
        }
        if (d[0] * d[1] * d[2] > 2147483647ll) {
          v.passthrough = 1;
          v.bits = ceil_log2_u64(C);
          slow = 1;
        } else {
          v.mul1 = (u32)div_b[0];
          v.mul2 = (u32)div_b[0] * (u32)div_b[1];
          const u64 cells = (u64)(u32)div_b[0] * (u32)div_b[1] * (u32)div_b[2];
          v.bits = cells >= (1ull << 31) ? 32u : ceil_log2_u64(cells);
        }
      }
      s.vfr = v;
      s.slow = slow;
      a.vf[f] = v;
      atomicMax(&a.ctl->fast_max_c, C);
    }
    __syncthreads();
    u32 V = 0;
    const bool slow0 = s.slow != 0;

    if (!slow0) {
      // ---- S1: load survivors, voxel keys
      u32 P2 = 32;
      while (P2 < C) P2 <<= 1;
      const VoxelFrame vfr = s.vfr;
      for (u32 i = tid; i < P2; i += kFrameThreads) {
        u64 k = 0xFFFFFFFFFFFFFFFFull;
        if (i < C) {
          const float4 p = a.pts[c0 + i];
          s.px[i] = p.x; s.py[i] = p.y; s.pz[i] = p.z; s.pw[i] = p.w;
          const i32 i0 = (i32)__fsub_rn(floorf(__fmul_rn(p.x, a.vk.inv[0])), (float)vfr.min_b[0]);
          const i32 i1 = (i32)__fsub_rn(floorf(__fmul_rn(p.y, a.vk.inv[1])), (float)vfr.min_b[1]);
          const i32 i2 = (i32)__fsub_rn(floorf(__fmul_rn(p.z, a.vk.inv[2])), (float)vfr.min_b[2]);
          const u32 idx = (u32)i0 + (u32)i1 * vfr.mul1 + (u32)i2 * vfr.mul2;
          k = ((u64)idx << 32) | (u64)i;
        }
        s.key[i] = k;
      }
      __syncthreads();
      // ---- S2: bitonic sort of (voxel idx, position): ascending idx, ascending position inside
      for (u32 k = 2; k <= P2; k <<= 1) {
        for (u32 j = k >> 1; j > 0; j >>= 1) {
          for (u32 t = tid; t < (P2 >> 1); t += kFrameThreads) {
            const u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
            const u32 ixj = i | j;
            const u64 x = s.key[i], y = s.key[ixj];
            const bool up = (i & k) == 0;
            if ((x > y) == up) {
              s.key[i] = y;
              s.key[ixj] = x;
            }
          }
          __syncthreads();
        }
      }
      // ---- S3: segment heads -> voxel ids
      const u32 per = (C + kFrameThreads - 1) / kFrameThreads;
      const u32 r0 = tid * per, r1 = min(r0 + per, C);
      u32 cnt = 0;
      for (u32 r = r0; r < r1; ++r)
        cnt += (r == 0 || (u32)(s.key[r] >> 32) != (u32)(s.key[r - 1] >> 32)) ? 1u : 0u;
      u32 total;
      u32 vid = block_excl_scan(cnt, s.wsum, total);
      V = total;
      if (V <= (u32)VMAX) {
        for (u32 r = r0; r < r1; ++r)
          if (r == 0 || (u32)(s.key[r] >> 32) != (u32)(s.key[r - 1] >> 32)) s.vstart[vid++] = r;
        if (tid == 0) s.vstart[V] = C;
      } else if (tid == 0) {
        s.slow = 1;
      }
      if (tid == 0) atomicMax(&a.ctl->fast_max_v, V);
      __syncthreads();
    }
    const bool slow = s.slow != 0;
    if (slow) V = 0;

    // ---- voxel offsets across frames (look-back in frame order)
    if (warp == 0) {
      const u32 e = lookback_exclusive(a.desc_v, f, V);
      if (lane == 0) {
        s.v_excl = e;
        a.v_off[f] = e;
        if (f + 1 == a.n_frames) {
          a.v_off[f + 1] = e + V;
          a.ctl->n_vox = e + V;
        }
        if (slow) atomicAdd(&a.ctl->fast_overflow, 1u);
      }
    }
    __syncthreads();
    const u32 v_excl = s.v_excl;

    u32 K = 0;
    if (!slow && V > 0) {
      // ---- S4: voxel centroids — sequential fp32 sums in record (= point) order
      for (u32 v = tid; v < V; v += kFrameThreads) {
        const u32 b = s.vstart[v], e = s.vstart[v + 1];
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        u32 cnt = e - b;
        for (u32 r = b; r < e; ++r) {
          const u32 i = (u32)s.key[r];
          sx = __fadd_rn(sx, s.px[i]);
          sy = __fadd_rn(sy, s.py[i]);
          sz = __fadd_rn(sz, s.pz[i]);
          si = __fadd_rn(si, s.pw[i]);
          if (a.pad_survives && a.src[c0 + i] == 0xFFFFFFFFu) {
            const u32 nf = a.uniform_n ? a.uniform_n : a.frame_n[f];
            cnt += (nf - a.gcount[f]) - 1u;
          }
        }
        const float c = (float)cnt;
        const float mx = __fdiv_rn(sx, c), my = __fdiv_rn(sy, c), mz = __fdiv_rn(sz, c);
        s.vx[v] = mx; s.vy[v] = my; s.vz[v] = mz;
        s.parent[v] = v;
        s.csize[v] = 0;
        if (a.tap_vox && v_excl + v < a.tap_vox_cap) a.tap_vox[v_excl + v] = make_float4(mx, my, mz, __fdiv_rn(si, c));
      }
      if (a.tap_keys) {
        for (u32 r = tid; r < C; r += kFrameThreads) {
          a.tap_keys[c0 + r] = (u32)(s.key[r] >> 32);
          a.tap_order[c0 + r] = c0 + (u32)s.key[r];
        }
      }
      __syncthreads();
      // ---- S5: connected components of "L2_Simple(i,j) < r2" — all pairs, warp per row.
      // A warp resolves a row cooperatively: roots of all hits, one warp-wide minimum, and a
      // CAS only for roots that still differ (rare once a component has formed).
      for (u32 i = warp + 1; i < V; i += kFrameThreads / 32) {
        const float xi = s.vx[i], yi = s.vy[i], zi = s.vz[i];
        for (u32 base = 0; base < i; base += 32) {
          const u32 j = base + lane;
          const bool hit = j < i && l2_simple(xi, yi, zi, s.vx[j], s.vy[j], s.vz[j]) < a.ck.r2;
          if (!__any_sync(kFull, hit)) continue;
          const u32 rj = hit ? smem_find(s.parent, j) : 0xFFFFFFFFu;
          const u32 ri = smem_find(s.parent, i);
          const u32 m = min(__reduce_min_sync(kFull, rj), ri);
          if (hit && rj != m && atomicCAS(&s.parent[rj], rj, m) != rj) smem_union(s.parent, rj, m);
          if (lane == 0 && ri != m && atomicCAS(&s.parent[ri], ri, m) != ri) smem_union(s.parent, ri, m);
        }
      }
      __syncthreads();
      // ---- S6: labels (root = min voxel index of the component) and component sizes
      for (u32 v = tid; v < V; v += kFrameThreads) s.label[v] = smem_find(s.parent, v);
      __syncthreads();
      for (u32 v = tid; v < V; v += kFrameThreads) {
        atomicAdd(&s.csize[s.label[v]], 1u);
        if (a.tap_labels) a.tap_labels[v_excl + v] = (i32)s.label[v];
      }
      __syncthreads();
      // ---- S7: kept roots in ascending order (ordered compaction over v)
      const u32 perv = (V + kFrameThreads - 1) / kFrameThreads;
      const u32 v0 = tid * perv, v1 = min(v0 + perv, V);
      u32 kc = 0, cc = 0;
      for (u32 v = v0; v < v1; ++v) {
        if (s.label[v] == v) {
          ++cc;
          const u32 sz = s.csize[v];
          if (sz >= a.ck.min_size && sz <= a.ck.max_size) ++kc;
        }
      }
      u32 ktotal, ctotal;
      u32 kpos = block_excl_scan(kc, s.wsum, ktotal);
      block_excl_scan(cc, s.wsum, ctotal);
      K = ktotal;
      for (u32 v = v0; v < v1; ++v) {
        if (s.label[v] == v) {
          const u32 sz = s.csize[v];
          if (sz >= a.ck.min_size && sz <= a.ck.max_size) s.vstart[kpos++] = v;  // vstart is free now
        }
      }
      if (tid == 0) s.n_comp = ctotal;
      __syncthreads();
    } else if (tid == 0) {
      s.n_comp = 0;
    }

    // ---- cluster offsets across frames
    if (warp == 0) {
      const u32 e = lookback_exclusive(a.desc_k, f, K);
      if (lane == 0) {
        s.k_excl = e;
        a.k_off[f] = e;
        a.kcount_f[f] = K;
        a.ncomp_f[f] = s.n_comp;
        if (f + 1 == a.n_frames) {
          a.k_off[f + 1] = e + K;
          a.ctl->n_clusters = e + K;
        }
        if ((u64)e + K > a.clusters_cap) atomicOr(&a.ctl->error, kErrVoxels);
      }
    }
    __syncthreads();
    // ---- S8/S9: centroid (src/cone_detection.cpp:261-273) and canonical rank, warp per cluster.
    // Lanes find the members 32 voxels at a time; the fp32 sums stay sequential in ascending
    // voxel index (every lane carries the same running sum).
    if (K > 0) {
      const u32 k_excl = s.k_excl;
      for (u32 k = warp; k < K; k += kFrameThreads / 32) {
        const u32 root = s.vstart[k];
        const u32 size = s.csize[root];
        float x = 0.0f, y = 0.0f;
        u32 seen = 0;
        for (u32 base = root; base < V && seen < size; base += 32) {
          const u32 v = base + lane;
          u32 members = __ballot_sync(kFull, v < V && s.label[v] == root);
          seen += __popc(members);
          while (members) {
            const u32 m = base + (u32)__ffs(members) - 1u;
            members &= members - 1;
            x = __fadd_rn(x, s.vx[m]);
            y = __fadd_rn(y, s.vy[m]);
          }
        }
        u32 rank = 0;  // size descending, then min index ascending (kept roots are ascending)
        for (u32 q = lane; q < K; q += 32) {
          const u32 sq = s.csize[s.vstart[q]];
          rank += (sq > size || (sq == size && q < k)) ? 1u : 0u;
        }
        rank = __reduce_add_sync(kFull, rank);
        if (lane == 0) {
          const float cnt = (float)(i32)size;
          ClusterRec o;
          o.x = __fdiv_rn(x, cnt);
          o.y = __fdiv_rn(y, cnt);
          o.size = size;
          o.min_index = root;
          if (k_excl + rank < a.clusters_cap) a.clusters[k_excl + rank] = o;
        }
      }
    }
  }
}

}  // namespace cp
