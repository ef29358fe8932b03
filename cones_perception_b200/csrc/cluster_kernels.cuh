// cluster_kernels.cuh — pcl::EuclideanClusterExtraction restated as connected components
// (reference call site: euclidan_cluster(), src/cone_detection.cpp:206-220; semantics:
// SURVEY.md Appendix A.5) plus the centroid loop (src/cone_detection.cpp:261-273, A.6).
//
//   cell_key      voxel centroid -> cell of a uniform grid with edge > cluster tolerance
//   (radix sort by cell key)
//   cell heads    -> cell start offsets;  hash_insert: cell key -> cell id (open addressing)
//   neighbour_union   27 adjacent cells, exact FLANN distance test, lock-free CAS union-find
//                     linking the larger root under the smaller (root = min index = label)
//   flatten_count     label = root, component sizes
//   (radix sort by label) -> component heads -> size filter + ordering key
//   (radix sort by frame, size desc) -> emit centroids in canonical cluster order
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace cp {

struct ClusterK {
  float r2;          // squared radius handed to FLANN: (float)((double)tol_f * tol_f)
  float inv_h;       // 1 / cell edge (edge = tolerance * 1.01)
  float origin;      // grid origin offset: coordinates + origin >= 0
  u32 nx;            // cells per axis
  u32 min_size, max_size;
  u32 frame_bits, size_bits;
};

__device__ __forceinline__ u32 cell_coord(float c, const ClusterK& k) {
  const float t = floorf((c + k.origin) * k.inv_h);
  i32 v = (i32)t;
  v = v < 0 ? 0 : v;
  v = v >= (i32)k.nx ? (i32)k.nx - 1 : v;
  return (u32)v;
}

__global__ void cluster_bits_kernel(Ctl* ctl, u32 csort_bits, u32 osort_bits) {
  ctl->csort_bits = csort_bits;
  ctl->lsort_bits = ceil_log2_u64(ctl->n_vox ? ctl->n_vox : 1);
  if (ctl->lsort_bits == 0) ctl->lsort_bits = 1;
  ctl->osort_bits = osort_bits;
}

__global__ void cell_key_kernel(const Ctl* __restrict__ ctl, ClusterK k, const float4* __restrict__ vox,
                                const u32* __restrict__ vox_frame, u64* __restrict__ keys,
                                u32* __restrict__ vals, u32* __restrict__ parent) {
  const u32 nv = ctl->n_vox;
  for (u32 v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    const float4 p = vox[v];
    const u64 cx = cell_coord(p.x, k), cy = cell_coord(p.y, k), cz = cell_coord(p.z, k);
    keys[v] = (((u64)vox_frame[v] * k.nx + cz) * k.nx + cy) * k.nx + cx;
    vals[v] = v;
    parent[v] = v;
  }
}

// live hash capacity = next pow2 >= 2 * n_cells; clear that prefix
__global__ void hash_setup_kernel(Ctl* ctl, u32 hash_cap) {
  u32 need = ctl->n_cells * 2u;
  u32 cap = 64;
  while (cap < need && cap < hash_cap) cap <<= 1;
  if (cap < need) atomicOr(&ctl->error, kErrHash);
  ctl->hash_mask = cap - 1;
}
constexpr u64 kHashEmpty = 0xFFFFFFFFFFFFFFFFull;
__global__ void hash_clear_kernel(const Ctl* __restrict__ ctl, u64* __restrict__ hkeys) {
  const u32 cap = ctl->hash_mask + 1;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x)
    hkeys[i] = kHashEmpty;
}
__device__ __forceinline__ u32 hash_u64(u64 k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (u32)k;
}
__global__ void hash_insert_kernel(const Ctl* __restrict__ ctl, const u64* keys_a, const u64* keys_b,
                                   const u32* __restrict__ cstart, u64* __restrict__ hkeys,
                                   u32* __restrict__ hvals) {
  const u32 nc = ctl->n_cells, mask = ctl->hash_mask;
  const u64* keys = sorted_in_b(ctl->csort_bits) ? keys_b : keys_a;
  for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
    const u64 key = keys[cstart[c]];
    u32 slot = hash_u64(key) & mask;
    for (u32 probe = 0; probe <= mask; ++probe) {
      const u64 old = atomicCAS((unsigned long long*)&hkeys[slot], (unsigned long long)kHashEmpty,
                                (unsigned long long)key);
      if (old == kHashEmpty || old == key) {
        hvals[slot] = c;
        break;
      }
      slot = (slot + 1) & mask;
    }
  }
}
__device__ __forceinline__ u32 hash_find(const u64* hkeys, const u32* hvals, u32 mask, u64 key) {
  u32 slot = hash_u64(key) & mask;
  for (u32 probe = 0; probe <= mask; ++probe) {
    const u64 k = hkeys[slot];
    if (k == key) return hvals[slot];
    if (k == kHashEmpty) return 0xFFFFFFFFu;
    slot = (slot + 1) & mask;
  }
  return 0xFFFFFFFFu;
}

// ---- lock-free union-find --------------------------------------------------------------
__device__ __forceinline__ u32 uf_find(u32* parent, u32 x) {
  // path halving; concurrent writers only ever replace a parent by an ancestor
  u32 p = ((volatile u32*)parent)[x];
  while (p != x) {
    const u32 gp = ((volatile u32*)parent)[p];
    if (gp != p) ((volatile u32*)parent)[x] = gp;
    x = p;
    p = gp;
  }
  return x;
}
__device__ __forceinline__ void uf_union(u32* parent, u32 a, u32 b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const u32 t = a;
      a = b;
      b = t;
    }
    // a > b: hang the larger root under the smaller one
    const u32 old = atomicCAS(&parent[a], a, b);
    if (old == a) return;
  }
}

// One warp per sorted voxel position.  Lanes 0..26 look up the 27 neighbour cells (one hash probe each, in
// parallel); then the warp walks the cells one after the other with its lanes strided over the candidates, so the
// index and parent loads of a crowded cell (a solid blob puts ~1000 voxels into one tolerance cell) are coalesced
// instead of 32 lanes each walking a list of their own.
__global__ void __launch_bounds__(256) neighbour_union_kernel(const Ctl* __restrict__ ctl, ClusterK k,
                                                              const u64* keys_a, const u64* keys_b, const u32* vals_a,
                                                              const u32* vals_b, const u32* __restrict__ cstart,
                                                              const u64* __restrict__ hkeys,
                                                              const u32* __restrict__ hvals,
                                                              const float4* __restrict__ vox, u32* __restrict__ parent,
                                                              Ctl* ctl_w) {
  const u32 nv = ctl->n_vox, nc = ctl->n_cells, mask = ctl->hash_mask;
  u32 n_visited = 0, n_tested = 0;   // statistics (cp_last_pairs)
  const bool inb = sorted_in_b(ctl->csort_bits);
  const u64* keys = inb ? keys_b : keys_a;
  const u32* vals = inb ? vals_b : vals_a;
  const u32 lane = threadIdx.x & 31u;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  for (u32 pos = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pos < nv; pos += nwarps) {
    const u64 ck = keys[pos];
    const u32 cx = (u32)(ck % k.nx), cy = (u32)((ck / k.nx) % k.nx), cz = (u32)((ck / ((u64)k.nx * k.nx)) % k.nx);
    u32 b = 0, e = 0;  // lane nb: candidate positions [b, e) of neighbour cell nb
    if (lane < 27u) {
      const i32 dx = (i32)(lane % 3u) - 1, dy = (i32)((lane / 3u) % 3u) - 1, dz = (i32)(lane / 9u) - 1;
      const i32 qx = (i32)cx + dx, qy = (i32)cy + dy, qz = (i32)cz + dz;
      if (qx >= 0 && qy >= 0 && qz >= 0 && qx < (i32)k.nx && qy < (i32)k.nx && qz < (i32)k.nx) {
        const u64 nk = (u64)((long long)ck + dx + (long long)dy * (long long)k.nx +
                             (long long)dz * (long long)k.nx * (long long)k.nx);
        const u32 c = hash_find(hkeys, hvals, mask, nk);
        if (c != 0xFFFFFFFFu) {
          b = cstart[c];
          // own cell: only earlier positions (each pair once); other cells: the whole cell, filtered by u < v
          e = (lane == 13u) ? pos : ((c + 1 < nc) ? cstart[c + 1] : nv);
        }
      }
    }
    const u32 v = vals[pos];
    const float4 p = vox[v];
    u32 rv = uf_find(parent, v);  // v's root, or (after other lanes' / warps' links) one of v's ancestors
    for (u32 nb = 0; nb < 27u; ++nb) {
      const u32 cb = __shfl_sync(kFull, b, nb), ce = __shfl_sync(kFull, e, nb);
      if (cb >= ce) continue;  // uniform over the warp
      for (u32 j = cb + lane; j < ce; j += 32u) {
        const u32 u = vals[j];
        if (nb != 13u && u > v) continue;  // each cross-cell pair is seen from both sides: test once
        ++n_visited;
        // u already hangs under v's root: the edge cannot change anything.  In a solid blob (thousands of
        // mutual neighbours per voxel) almost every candidate leaves here after one 4-byte load.
        if (((volatile u32*)parent)[u] == rv) continue;
        const float4 q = vox[u];
        ++n_tested;
        // FLANN: dist = L2_Simple(query, point); accepted iff dist < radius (strict)
        if (l2_simple(p.x, p.y, p.z, q.x, q.y, q.z) < k.r2) {
          u32 ru = uf_find(parent, u);
          if (ru != u) ((volatile u32*)parent)[u] = ru;  // u is not a root: point it at its root for later visitors
          while (ru != rv) {  // larger root under the smaller; parents always have smaller indices
            const u32 hi = rv > ru ? rv : ru, sm = rv > ru ? ru : rv;
            const u32 old = atomicCAS(&parent[hi], hi, sm);
            if (old == hi) {
              rv = sm;
              break;
            }
            if (hi == rv) rv = uf_find(parent, old);
            else ru = uf_find(parent, old);
          }
        }
      }
      __syncwarp();
      rv = __reduce_min_sync(kFull, rv);  // every lane holds v or an ancestor of v: the smallest is the highest
    }
    if (lane == 0 && rv != v) ((volatile u32*)parent)[v] = rv;
  }
  n_visited = __reduce_add_sync(kFull, n_visited);
  n_tested = __reduce_add_sync(kFull, n_tested);
  if (lane == 0 && n_visited) {
    atomicAdd(&ctl_w->pairs_visited, (unsigned long long)n_visited);
    atomicAdd(&ctl_w->pairs_tested, (unsigned long long)n_tested);
  }
}

// label = root (= min voxel index of the component); keys for the label sort
__global__ void flatten_kernel(const Ctl* __restrict__ ctl, u32* __restrict__ parent, u32* __restrict__ label,
                               u64* __restrict__ keys, u32* __restrict__ vals) {
  const u32 nv = ctl->n_vox;
  for (u32 v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    const u32 r = uf_find(parent, v);
    label[v] = r;
    keys[v] = r;
    vals[v] = v;
  }
}

// one thread per component (segment of the label-sorted voxel list)
__global__ void component_kernel(const Ctl* __restrict__ ctl, ClusterK k, const u64* keys_a, const u64* keys_b,
                                 const u32* __restrict__ comp_start, const u32* __restrict__ vox_frame,
                                 u32* __restrict__ ncomp_f, u32* __restrict__ kcount_f,
                                 u64* __restrict__ okeys, u32* __restrict__ ovals, Ctl* ctl_w) {
  const u32 ncomp = ctl->n_comp, nv = ctl->n_vox;
  const u64* keys = sorted_in_b(ctl->lsort_bits) ? keys_b : keys_a;
  for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < ncomp; c += gridDim.x * blockDim.x) {
    const u32 b = comp_start[c];
    const u32 e = (c + 1 < ncomp) ? comp_start[c + 1] : nv;
    const u32 size = e - b;
    const u32 root = (u32)keys[b];
    const u32 f = vox_frame[root];
    atomicAdd(&ncomp_f[f], 1u);
    const bool kept = size >= k.min_size && size <= k.max_size;
    u64 ok;
    if (kept) {
      atomicAdd(&kcount_f[f], 1u);
      atomicAdd(&ctl_w->n_clusters, 1u);
      ok = ((u64)f << k.size_bits) | (u64)(k.max_size - size);  // size descending inside the frame
    } else {
      ok = 1ull << (k.frame_bits + k.size_bits);  // dropped components sort behind every kept one
    }
    okeys[c] = ok;
    ovals[c] = c;
  }
}

// exclusive scan of the per-frame cluster counts (one CTA; F is small)
__global__ void __launch_bounds__(1024) frame_scan_kernel(u32 n_frames, const u32* __restrict__ cnt,
                                                          u32* __restrict__ off) {
  __shared__ u32 wsum[32];
  __shared__ u32 carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  for (u32 c0 = 0; c0 < n_frames; c0 += blockDim.x) {
    const u32 i = c0 + threadIdx.x;
    const u32 v = i < n_frames ? cnt[i] : 0u;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const u32 w = wsum[lane];
      u32 winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(kFull, winc, o);
        if (lane >= o) winc += t;
      }
      wsum[lane] = winc - w;
    }
    __syncthreads();
    const u32 excl = carry_s + wsum[warp] + inc - v;
    if (i < n_frames) off[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[n_frames] = carry_s;
}

struct ClusterRec {
  float x, y;
  u32 size, min_index;
};

// one thread per kept cluster, in canonical order (frame, size desc, min index asc)
__global__ void emit_clusters_kernel(const Ctl* __restrict__ ctl, const u32* ovals_a, const u32* ovals_b,
                                     const u64* lkeys_a, const u64* lkeys_b, const u32* lvals_a,
                                     const u32* lvals_b, const u32* __restrict__ comp_start,
                                     const float4* __restrict__ vox, const u32* __restrict__ vox_frame,
                                     const u32* __restrict__ v_off, ClusterRec* __restrict__ out, u32 out_cap,
                                     Ctl* ctl_w) {
  const u32 nk = ctl->n_clusters, ncomp = ctl->n_comp, nv = ctl->n_vox;
  const u32* ovals = sorted_in_b(ctl->osort_bits) ? ovals_b : ovals_a;
  const bool lb = sorted_in_b(ctl->lsort_bits);
  const u64* lkeys = lb ? lkeys_b : lkeys_a;
  const u32* lvals = lb ? lvals_b : lvals_a;
  for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < nk; r += gridDim.x * blockDim.x) {
    if (r >= out_cap) {
      atomicOr(&ctl_w->error, kErrVoxels);
      continue;
    }
    const u32 c = ovals[r];
    const u32 b = comp_start[c];
    const u32 e = (c + 1 < ncomp) ? comp_start[c + 1] : nv;
    const u32 root = (u32)lkeys[b];
    // src/cone_detection.cpp:264-268: float x, y accumulated over ascending voxel indices
    float x = 0.0f, y = 0.0f;
    for (u32 j = b; j < e; ++j) {
      const float4 p = vox[lvals[j]];
      x = __fadd_rn(x, p.x);
      y = __fadd_rn(y, p.y);
    }
    const float cnt = (float)(i32)(e - b);
    ClusterRec o;
    o.x = __fdiv_rn(x, cnt);
    o.y = __fdiv_rn(y, cnt);
    o.size = e - b;
    o.min_index = root - v_off[vox_frame[root]];
    out[r] = o;
  }
}

// frame-local canonical labels for the parity tap
__global__ void local_labels_kernel(const Ctl* __restrict__ ctl, const u32* __restrict__ label,
                                    const u32* __restrict__ vox_frame, const u32* __restrict__ v_off,
                                    i32* __restrict__ out) {
  const u32 nv = ctl->n_vox;
  for (u32 v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x)
    out[v] = (i32)(label[v] - v_off[vox_frame[v]]);
}

}  // namespace cp
