// cluster_kernels.cuh — pcl::EuclideanClusterExtraction restated as connected components
// (reference call site: euclidan_cluster(), src/cone_detection.cpp:206-220; semantics:
// SURVEY.md Appendix A.5) plus the centroid loop (src/cone_detection.cpp:261-273, A.6).
//
//   cell_key      voxel centroid -> cell of a uniform grid with edge 0.505 x cluster tolerance; also the sort
//                 widths of the stage, the cell hash's live size and clearing, the cell sort's digit histograms
//   (radix sort by cell key)
//   cell heads    -> cell start offsets; every head enters cell key -> cell id into the hash (open addressing)
//   cell_union    one warp per occupied cell.  The cell's diagonal is shorter than the tolerance, so its voxels
//                 are one component without a single distance test; the 62 "forward" cells of the 5x5x5
//                 neighbourhood are then joined cell to cell: already in one tree -> skipped after two finds,
//                 otherwise voxel pairs are tested (exact FLANN distance) until the first hit.  Lock-free CAS
//                 union-find linking the larger root under the smaller (root = min index = label).
//   flatten       label = root; keys + digit histograms of the label sort
//   (radix sort by label) -> component heads -> component: size filter + ordering key (+ order sort histograms)
//   (radix sort by frame, size desc) -> emit centroids in canonical cluster order; its CTA 0 also writes the
//                 cluster offsets per frame and the per-frame counters
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"
#include "voxel_kernels.cuh"

namespace cp {

struct ClusterK {
  float r2;          // squared radius handed to FLANN: (float)((double)tol_f * tol_f)
  float inv_h;       // 1 / sweep cell edge of the per-frame kernel (edge = tolerance * 1.01)
  float inv_g;       // 1 / cell edge of the general path's grid (edge = tolerance * 0.505: diagonal < tolerance)
  float origin;      // grid origin offset: coordinates + origin >= 0
  u32 nx;            // cells per axis of the general path's grid
  u32 min_size, max_size;
  u32 frame_bits, size_bits;
};

// `outside` is raised when the coordinate lies beyond the grid (cannot happen behind the distance crop: the grid
// spans distance_treshold_max plus a margin); the cell arguments of cell_union_kernel do not hold for a clamped cell
__device__ __forceinline__ u32 cell_coord(float c, const ClusterK& k, bool& outside) {
  const float t = floorf((c + k.origin) * k.inv_g);
  i32 v = (i32)t;
  if (!(t >= 0.0f) || !(t < (float)k.nx)) outside = true;
  v = v < 0 ? 0 : v;
  v = v >= (i32)k.nx ? (i32)k.nx - 1 : v;
  return (u32)v;
}

// live hash capacity: next power of two >= 2 x the voxel count (an upper bound of the occupied cells, known before
// the cells are); every thread can work it out for itself
__device__ __forceinline__ u32 hash_capacity(u32 n_vox, u32 hash_cap) {
  const u32 need = n_vox * 2u;
  u32 cap = 64;
  while (cap < need && cap < hash_cap) cap <<= 1;
  return cap;
}

// one thread per voxel: key of its neighbour-grid cell (prefixed by the frame), union-find parent = itself.  The
// kernel also settles what used to be kernels of their own: the key widths of the three cluster-stage sorts, the
// live size of the cell hash and its clearing, and the digit histograms of the cell sort (sort_feed_*).
__global__ void __launch_bounds__(256) cell_key_kernel(Ctl* ctl, ClusterK k, const float4* __restrict__ vox,
                                                       const u32* __restrict__ vox_frame, u64* __restrict__ keys,
                                                       u32* __restrict__ vals, u32* __restrict__ parent,
                                                       u32 csort_bits, u32 osort_bits, u32 hash_cap,
                                                       u64* __restrict__ hkeys, u32* sort_hdr, u32* sort_state) {
  __shared__ SortFeedSmem feed;
  const u32 nv = ctl->n_vox;
  const u32 hcap = hash_capacity(nv, hash_cap);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->csort_bits = csort_bits;
    u32 lb = ceil_log2_u64(nv ? nv : 1);
    ctl->lsort_bits = lb ? lb : 1u;
    ctl->osort_bits = osort_bits;
    ctl->hash_mask = hcap - 1;
    if (hcap < nv * 2u) atomicOr(&ctl->error, kErrHash);
  }
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < hcap; i += stride) hkeys[i] = kHashEmpty;
  const u32 passes = sort_feed_passes(csort_bits, nv);
  sort_feed_begin(feed, passes);
  bool outside = false;
  for (u32 base = blockIdx.x * blockDim.x; base < nv; base += stride) {   // warp-uniform bounds (sort_feed_key)
    const u32 v = base + threadIdx.x;
    const bool valid = v < nv;
    u64 key = 0;
    if (valid) {
      const float4 p = vox[v];
      const u64 cx = cell_coord(p.x, k, outside), cy = cell_coord(p.y, k, outside), cz = cell_coord(p.z, k, outside);
      key = (((u64)vox_frame[v] * k.nx + cz) * k.nx + cy) * k.nx + cx;
      keys[v] = key;
      vals[v] = v;
      parent[v] = v;
    }
    sort_feed_key(feed, key, valid, passes);
  }
  if (outside) atomicOr(&ctl->error, kErrInternal);   // never silently: surfaces as an error at cp_sync
  sort_feed_flush(feed, sort_hdr, sort_state, passes, nv);
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

// The 62 cells of the 5x5x5 neighbourhood that come after the centre in (dz, dy, dx) order — every pair of cells
// is visited from one side only — nearest first, so that face neighbours (which almost always connect at the
// first pairs tested) are joined before the far corners are looked at and usually found connected already.
__device__ const signed char kForwardCells[62][3] = {
    {1,0,0}, {0,1,0}, {0,0,1}, {-1,1,0}, {1,1,0}, {0,-1,1}, {-1,0,1}, {1,0,1}, {0,1,1}, {-1,-1,1}, {1,-1,1}, {-1,1,1},
    {1,1,1}, {2,0,0}, {0,2,0}, {0,0,2}, {-2,1,0}, {2,1,0}, {-1,2,0}, {1,2,0}, {0,-2,1}, {-2,0,1}, {2,0,1}, {0,2,1},
    {0,-1,2}, {-1,0,2}, {1,0,2}, {0,1,2}, {-1,-2,1}, {1,-2,1}, {-2,-1,1}, {2,-1,1}, {-2,1,1}, {2,1,1}, {-1,2,1},
    {1,2,1}, {-1,-1,2}, {1,-1,2}, {-1,1,2}, {1,1,2}, {-2,2,0}, {2,2,0}, {0,-2,2}, {-2,0,2}, {2,0,2}, {0,2,2},
    {-2,-2,1}, {2,-2,1}, {-2,2,1}, {2,2,1}, {-1,-2,2}, {1,-2,2}, {-2,-1,2}, {2,-1,2}, {-2,1,2}, {2,1,2}, {-1,2,2},
    {1,2,2}, {-2,-2,2}, {2,-2,2}, {-2,2,2}, {2,2,2}};

// One warp per occupied cell of the grid with edge g = 0.505 x tolerance.
//  * 3 g^2 = 0.765 tol^2 < r2: any two voxels of one cell pass FLANN's test (with a 23 % margin against fp32
//    rounding), so the cell is linked into one tree without evaluating a distance.  Dense data is where the
//    pair-by-pair version drowned: a solid 1 m^3 blob gives every voxel ~4000 neighbours (62 M pair visits per
//    frame of config 5); a cell holds ~125 of them.
//  * two voxels within the tolerance differ by fewer than tol / g = 1.98 cells per axis, i.e. by at most 2 cell
//    indices: all edges lie inside the 5x5x5 neighbourhood.  For a neighbour cell that already hangs under the
//    same root nothing is tested; otherwise the warp's lanes walk the voxel pairs of the two cells until one
//    passes the exact test (strict <, FLANN L2_Simple) and links the two trees.  If no pair passes, the cells are
//    not directly connected — an edge between them would have been found, every pair is looked at.
// Components and labels (root = smallest voxel index) are those of the pair-by-pair graph.
__global__ void __launch_bounds__(256) cell_union_kernel(const Ctl* __restrict__ ctl, ClusterK k,
                                                         const u64* keys_a, const u64* keys_b, const u32* vals_a,
                                                         const u32* vals_b, const u32* __restrict__ cstart,
                                                         const u64* __restrict__ hkeys,
                                                         const u32* __restrict__ hvals,
                                                         const float4* __restrict__ vox, u32* __restrict__ parent,
                                                         Ctl* ctl_w) {
  const u32 nv = ctl->n_vox, nc = ctl->n_cells, mask = ctl->hash_mask;
  const bool inb = sorted_in_b(ctl->csort_bits);
  const u64* keys = inb ? keys_b : keys_a;
  const u32* vals = inb ? vals_b : vals_a;
  const u32 lane = threadIdx.x & 31u;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  unsigned long long n_visited = 0, n_tested = 0;   // statistics (cp_last_pairs)
  for (u32 c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < nc; c += nwarps) {
    const u32 b = cstart[c], e = (c + 1 < nc) ? cstart[c + 1] : nv;
    const u64 ck = keys[b];
    const i32 cx = (i32)(ck % k.nx), cy = (i32)((ck / k.nx) % k.nx), cz = (i32)((ck / ((u64)k.nx * k.nx)) % k.nx);
    // ---- the cell itself: one tree under its smallest voxel index, no distance tests
    u32 m = 0xFFFFFFFFu;
    for (u32 j = b + lane; j < e; j += 32u) m = min(m, vals[j]);
    m = __reduce_min_sync(kFull, m);
    for (u32 j = b + lane; j < e; j += 32u) {
      const u32 v = vals[j];
      if (v != m) uf_union(parent, v, m);
    }
    __syncwarp();
    // ---- the forward neighbour cells, nearest first
    for (u32 round = 0; round < 2u; ++round) {
      const u32 q = round * 32u + lane;
      u32 nb_b = 0, nb_e = 0;
      if (q < 62u) {
        const i32 dx = kForwardCells[q][0], dy = kForwardCells[q][1], dz = kForwardCells[q][2];
        const i32 qx = cx + dx, qy = cy + dy, qz = cz + dz;
        if (qx >= 0 && qy >= 0 && qz >= 0 && qx < (i32)k.nx && qy < (i32)k.nx && qz < (i32)k.nx) {
          const u64 nk = (u64)((long long)ck + dx + (long long)dy * (long long)k.nx +
                               (long long)dz * (long long)k.nx * (long long)k.nx);
          const u32 n = hash_find(hkeys, hvals, mask, nk);
          if (n != 0xFFFFFFFFu) {
            nb_b = cstart[n];
            nb_e = (n + 1 < nc) ? cstart[n + 1] : nv;
          }
        }
      }
      u32 found = __ballot_sync(kFull, nb_e > nb_b);
      while (found) {
        const int src = __ffs(found) - 1;
        found &= found - 1;
        const u32 ob = __shfl_sync(kFull, nb_b, src), oe = __shfl_sync(kFull, nb_e, src);
        // already one tree (through this or any other chain of cells): nothing to test
        u32 same = 0;
        if (lane == 0) same = uf_find(parent, m) == uf_find(parent, vals[ob]) ? 1u : 0u;
        if (__shfl_sync(kFull, same, 0)) continue;
        // Every voxel pair of the two cells until the first hit.  A lane keeps one voxel of the cell, a chunk of 32
        // neighbour voxels is fetched once (one per lane) and handed round by shuffles, so a pair costs arithmetic,
        // not a dependent index -> voxel load pair (two dense cells that are close but not connected — all 64 x 64
        // pairs fail — would otherwise be 128 memory round trips for one warp).
        const u32 nb = oe - ob, na = e - b;
        bool linked = false;
        for (u32 a0 = 0; a0 < na && !linked; a0 += 32u) {
          const bool pv = a0 + lane < na;
          u32 vi = 0;
          float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
          if (pv) {
            vi = vals[b + a0 + lane];
            p = vox[vi];
          }
          for (u32 c0 = 0; c0 < nb && !linked; c0 += 32u) {
            u32 vj = 0;
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + lane < nb) {
              vj = vals[ob + c0 + lane];
              r = vox[vj];
            }
            const u32 cnt = nb - c0 < 32u ? nb - c0 : 32u;
            for (u32 j = 0; j < cnt; ++j) {
              const float rx = __shfl_sync(kFull, r.x, j), ry = __shfl_sync(kFull, r.y, j),
                          rz = __shfl_sync(kFull, r.z, j);
              bool hit = false;
              if (pv) {
                ++n_visited;
                ++n_tested;
                // FLANN: dist = L2_Simple(query, point); accepted iff dist < radius (strict)
                hit = l2_simple(p.x, p.y, p.z, rx, ry, rz) < k.r2;
              }
              const u32 hits = __ballot_sync(kFull, hit);
              if (hits) {
                const u32 vjj = __shfl_sync(kFull, vj, j);
                if ((int)lane == __ffs(hits) - 1) uf_union(parent, vi, vjj);
                linked = true;
                break;
              }
            }
          }
        }
        __syncwarp();
      }
    }
  }
  for (int o = 16; o; o >>= 1) {
    n_visited += __shfl_xor_sync(kFull, n_visited, o);
    n_tested += __shfl_xor_sync(kFull, n_tested, o);
  }
  if (lane == 0 && n_visited) {
    atomicAdd(&ctl_w->pairs_visited, n_visited);
    atomicAdd(&ctl_w->pairs_tested, n_tested);
  }
}

// label = root (= min voxel index of the component); keys and digit histograms of the label sort
__global__ void __launch_bounds__(256) flatten_kernel(const Ctl* __restrict__ ctl, u32* __restrict__ parent,
                                                      u32* __restrict__ label, u64* __restrict__ keys,
                                                      u32* __restrict__ vals, u32* sort_hdr, u32* sort_state) {
  __shared__ SortFeedSmem feed;
  const u32 nv = ctl->n_vox;
  const u32 passes = sort_feed_passes(ctl->lsort_bits, nv);
  sort_feed_begin(feed, passes);
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 base = blockIdx.x * blockDim.x; base < nv; base += stride) {
    const u32 v = base + threadIdx.x;
    const bool valid = v < nv;
    u64 key = 0;
    if (valid) {
      const u32 r = uf_find(parent, v);
      label[v] = r;
      key = r;
      keys[v] = key;
      vals[v] = v;
    }
    sort_feed_key(feed, key, valid, passes);
  }
  sort_feed_flush(feed, sort_hdr, sort_state, passes, nv);
}

// one thread per component (segment of the label-sorted voxel list)
__global__ void __launch_bounds__(256) component_kernel(const Ctl* __restrict__ ctl, ClusterK k, const u64* keys_a,
                                                        const u64* keys_b, const u32* __restrict__ comp_start,
                                                        const u32* __restrict__ vox_frame,
                                                        u32* __restrict__ ncomp_f, u32* __restrict__ kcount_f,
                                                        u64* __restrict__ okeys, u32* __restrict__ ovals, Ctl* ctl_w,
                                                        u32* sort_hdr, u32* sort_state) {
  __shared__ SortFeedSmem feed;
  const u32 ncomp = ctl->n_comp, nv = ctl->n_vox;
  const u64* keys = sorted_in_b(ctl->lsort_bits) ? keys_b : keys_a;
  const u32 passes = sort_feed_passes(ctl->osort_bits, ncomp);
  sort_feed_begin(feed, passes);
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 base = blockIdx.x * blockDim.x; base < ncomp; base += stride) {
    const u32 c = base + threadIdx.x;
    const bool valid = c < ncomp;
    u64 ok = 0;
    if (valid) {
      const u32 b = comp_start[c];
      const u32 e = (c + 1 < ncomp) ? comp_start[c + 1] : nv;
      const u32 size = e - b;
      const u32 root = (u32)keys[b];
      const u32 f = vox_frame[root];
      atomicAdd(&ncomp_f[f], 1u);
      const bool kept = size >= k.min_size && size <= k.max_size;
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
    sort_feed_key(feed, ok, valid, passes);
  }
  sort_feed_flush(feed, sort_hdr, sort_state, passes, ncomp);
}

// exclusive scan of the per-frame cluster counts by one CTA of 256 threads (F is small)
__device__ __forceinline__ void frame_scan_block(u32 n_frames, const u32* __restrict__ cnt, u32* __restrict__ off) {
  __shared__ u32 wsum[8];
  __shared__ u32 carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  for (u32 c0 = 0; c0 < n_frames; c0 += 256u) {
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
    u32 woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const u32 t = wsum[w];
      if (w < warp) woff += t;
      total += t;
    }
    const u32 excl = carry_s + woff + inc - v;
    if (i < n_frames) off[i] = excl;
    __syncthreads();
    if (threadIdx.x == 0) carry_s += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[n_frames] = carry_s;
}

// what the last kernel of the general back half also settles (its CTA 0): cluster offsets per frame and the
// per-frame counters in the cp_frame_counters layout
struct FinishArgs {
  u32 n_frames;
  const u32* frame_n;
  u32 uniform_n;
  const u32* c_off;
  const u32* ncomp_f;
  const u32* kcount_f;
  const VoxelFrame* vf;
  const u32* gcount;
  int counted_ground;
  u32* k_off;   // [F+1]
  u32* fc;      // [F][8]
};

struct ClusterRec {
  float x, y;
  u32 size, min_index;
};

// one warp per kept cluster, in canonical order (frame, size desc, min index asc): the lanes fetch 32 members at
// a time and every lane adds them in ascending voxel index, so the fp32 sums stay sequential
// (src/cone_detection.cpp:264-268) without one thread chasing index -> voxel loads one by one
__global__ void __launch_bounds__(256) emit_clusters_kernel(const Ctl* __restrict__ ctl, const u32* ovals_a, const u32* ovals_b,
                                     const u64* lkeys_a, const u64* lkeys_b, const u32* lvals_a,
                                     const u32* lvals_b, const u32* __restrict__ comp_start,
                                     const float4* __restrict__ vox, const u32* __restrict__ vox_frame,
                                     const u32* __restrict__ v_off, ClusterRec* __restrict__ out, u32 out_cap,
                                     Ctl* ctl_w, FinishArgs fin) {
  if (blockIdx.x == 0) {
    frame_scan_block(fin.n_frames, fin.kcount_f, fin.k_off);
    for (u32 f = threadIdx.x; f < fin.n_frames; f += blockDim.x) {
      u32* o = fin.fc + (u64)f * 8;
      o[0] = fin.uniform_n ? fin.uniform_n : fin.frame_n[f];
      o[1] = fin.counted_ground ? fin.gcount[f] : o[0];   // G (= N without ground removal)
      o[2] = fin.c_off[f + 1] - fin.c_off[f];
      o[3] = v_off[f + 1] - v_off[f];
      o[4] = fin.ncomp_f[f];
      o[5] = fin.kcount_f[f];
      o[6] = fin.vf[f].bits;
      o[7] = fin.vf[f].passthrough;
    }
  }
  const u32 nk = ctl->n_clusters, ncomp = ctl->n_comp, nv = ctl->n_vox;
  const u32* ovals = sorted_in_b(ctl->osort_bits) ? ovals_b : ovals_a;
  const bool lb = sorted_in_b(ctl->lsort_bits);
  const u64* lkeys = lb ? lkeys_b : lkeys_a;
  const u32* lvals = lb ? lvals_b : lvals_a;
  const u32 lane = threadIdx.x & 31u;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  for (u32 r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nk; r += nwarps) {
    if (r >= out_cap) {
      if (lane == 0) atomicOr(&ctl_w->error, kErrVoxels);
      continue;
    }
    const u32 c = ovals[r];
    const u32 b = comp_start[c];
    const u32 e = (c + 1 < ncomp) ? comp_start[c + 1] : nv;
    const u32 root = (u32)lkeys[b];
    float x = 0.0f, y = 0.0f;
    // four rounds of members (128) are fetched before the first is added: a cone is one or two rounds, so the
    // index -> voxel load chain is paid once per cluster, not once per 32 members
    for (u32 j0 = b; j0 < e; j0 += 128u) {
      float px[4], py[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const u32 j = j0 + u * 32u + lane;
        px[u] = 0.f;
        py[u] = 0.f;
        if (j < e) {
          const float4 p = vox[lvals[j]];
          px[u] = p.x;
          py[u] = p.y;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const u32 r0 = j0 + u * 32u;
        const u32 m = r0 >= e ? 0u : (e - r0 < 32u ? e - r0 : 32u);
        for (u32 q = 0; q < m; ++q) {
          x = __fadd_rn(x, __shfl_sync(kFull, px[u], q));
          y = __fadd_rn(y, __shfl_sync(kFull, py[u], q));
        }
      }
    }
    if (lane == 0) {
      const float cnt = (float)(i32)(e - b);
      ClusterRec o;
      o.x = __fdiv_rn(x, cnt);
      o.y = __fdiv_rn(y, cnt);
      o.size = e - b;
      o.min_index = root - v_off[vox_frame[root]];
      out[r] = o;
    }
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
