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
//
// The kernel also gathers its frame's survivors straight from the keep mask (ordered, by word
// popcounts) and reduces their bounding box in shared memory, so the fast path needs neither
// the tile scan nor the global gather pass nor the survivor arrays.
#pragma once
#include "cluster_kernels.cuh"
#include "common.cuh"
#include "stream_kernels.cuh"
#include "voxel_kernels.cuh"

namespace cp {

constexpr int kFrameThreadsMax = 512;

#ifdef CP_PHASE_CLOCKS
// instrumented build (-DCP_PHASE_CLOCKS): thread 0 of the CTA that owns frame 0 stamps every phase
// and prints the cycle counts when the frame is done (one printf, so the phases are not perturbed)
#define CP_PHASE(name)                                                          \
  do {                                                                          \
    __syncthreads();                                                            \
    if (tid == 0 && f == 0 && clk_n < 16) {                                     \
      clk_name[clk_n] = name;                                                   \
      clk_t[clk_n++] = clock64();                                               \
    }                                                                           \
  } while (0)
#else
#define CP_PHASE(name) do { } while (0)
#endif

struct FrameArgs {
  u32 n_frames;
  const uint8_t* in;         // raw input points (the kernel gathers its frame's survivors itself)
  Layout layout;
  Geom geom;
  // keep bits: either read from `mask` (written by keep_mask_kernel), or — mask == NULL — evaluated by the
  // frame's own CTA (pass 2 of the ground filter + the crop, fused in front of the back half)
  const u32* mask;
  const u32* rowmax;         // pass 1's per-row maxima (NULL: every row must be evaluated)
  const u32* low_key;        // [F][32] per-sector minima of pass 1 (ground removal on)
  CropK crop;
  GroundK gk;
  u32* rows_loaded;          // statistics: rows the fused pass 2 read
  const u32* c_off;          // [F+1] survivor offsets — only with taps (NULL otherwise)
  u32* gcount;               // [F] ground survivors per frame (written here when mask == NULL)
  int pad_survives;
  VoxelK vk;
  ClusterK ck;
  // outputs
  VoxelFrame* vf;            // [F]
  u32* v_off;                // [F+1]
  u32* k_off;                // [F+1]
  u32* ncomp_f;              // [F]
  u32* kcount_f;             // [F]
  u32* ncrop_f;              // [F] survivors per frame
  ClusterRec* slots;         // [F][VMAX] per-frame result slots, canonical order inside a frame
  ClusterRec* direct_out;    // single-frame batches: the packed result list itself (no pack kernel); else NULL
  u32 direct_cap;
  u32* nvox_f;               // [F] voxels per frame
  u32* fc;                   // [F][8] cp_frame_counters records
  int counted_ground;
  u64* desc_v;               // frame descriptors for the voxel offsets
  Ctl* ctl;
  u32* ticket;               // frame ticket (zeroed before the launch)
  // taps (NULL when off)
  float4* tap_vox;
  u32 tap_vox_cap;
  u32* tap_keys;
  u32* tap_order;
  i32* tap_labels;
};

template <int CMAX, int VMAX, int T>
struct FrameSmem {
  float px[CMAX], py[CMAX], pz[CMAX], pw[CMAX];
  u32 k0[CMAX], k1[CMAX];              // voxel idx, ping-pong buffers of the radix sort
  unsigned short v0[CMAX], v1[CMAX];   // survivor position, ping-pong
  union {
    struct {                           // scratch of the radix sort (dead before the voxel arrays live)
      u32 whist[T / 32][256];
      u32 gbase[256];
    } sort;
    struct {
      float vx[VMAX], vy[VMAX], vz[VMAX];
      u32 parent[VMAX];
      u32 label[VMAX];                 // also: per-cell fill counters of the sweep
      u32 csize[VMAX];
      unsigned short vcell[VMAX];      // sweep cell of each voxel
      unsigned short perm[VMAX];       // voxels ordered by sweep cell
    } vox;
  } u;
  u32 vstart[VMAX + 1];                // voxel -> first sorted record; then sweep cell starts; then kept roots
  u32 wsum[T / 32];
  VoxelFrame vfr;
  u32 frame, slow, v_excl, n_comp;
  u32 gkept, nlive;                    // fused pass 2: ground survivors / live rows of the frame
  float thr[kSectStride];              // ground thresholds of the frame; [31] their minimum, [30] their maximum
  u32 bbox[8];
  u32 sw_axis, sw_ncell;
  float sw_min, sw_inv;
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
template <int T>
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
  for (int w = 0; w < T / 32; ++w) {
    const u32 t = wsum[w];
    if (w < warp) woff += t;
    total += t;
  }
  return woff + inc - v;
}

// The whole back half of frame f by one CTA of T threads (see the file header); called by frame_backend_kernel
// for batches and by single_frame_kernel (single_frame.cuh) for a node's single frame.
template <int CMAX, int VMAX, int MODE, int T>
__device__ __forceinline__ void frame_process(const FrameArgs& a, FrameSmem<CMAX, VMAX, T>& s, const u32 f) {
  constexpr int kFrameThreads = T;
  const u32 tid = threadIdx.x;
  const int lane = lane_id(), warp = tid >> 5;
  {
#ifdef CP_PHASE_CLOCKS
    const char* clk_name[16];
    long long clk_t[16];
    int clk_n = 0;
    const long long clk_0 = clock64();
#endif
    // frame geometry: tiles [tile0, tile0 + ntiles), points [first, first + n)
    u32 tile0, ntiles, npts;
    u64 first;
    if (a.geom.uniform_n) {
      tile0 = f * a.geom.tpf;
      ntiles = a.geom.tpf;
      npts = a.geom.uniform_n;
      first = (u64)f * a.geom.uniform_n;
    } else {
      tile0 = a.geom.frame_tile0[f];
      ntiles = (f + 1 < a.n_frames ? a.geom.frame_tile0[f + 1] : a.geom.n_tiles) - tile0;
      npts = a.geom.frame_n[f];
      first = a.geom.frame_off[f];
    }
    const u32 c0 = a.c_off ? a.c_off[f] : 0u;

    // ---- S0: keep bits of the frame, then the ordered gather of its survivors + bounding box.
    // The frame is walked in chunks of CH = 4 * CMAX rows (one row = 32 consecutive points = one mask word);
    // the chunk's mask words live in shared memory, aliased over the point arrays that are filled afterwards.
    if (tid < 8) s.bbox[tid] = tid < 4 ? 0xFFFFFFFFu : 0u;
    if (tid == 0) {
      s.slow = 0;
      s.gkept = 0;
    }
    const bool fused = a.mask == nullptr;
    const bool do_ground = fused && a.gk.do_ground != 0;   // (with a ready-made mask the ground verdicts are in it)
    if (warp == 0) {
      // :75  p.z < low + 0.1 in double  <=>  z < roundup_to_float((double)low + 0.1)
      float t = 0.0f;
      if (do_ground && lane < kNSect) {
        t = __double2float_ru((double)ord2f(__ldcg(&a.low_key[f * kSectStride + lane])) + 0.1);
        s.thr[lane] = t;
      }
      float mn = (do_ground && lane < kNSect) ? t : __int_as_float(0x7f800000);
      float mx = (do_ground && lane < kNSect) ? t : -__int_as_float(0x7f800000);
#pragma unroll
      for (int o2 = 16; o2; o2 >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(kFull, mn, o2));
        mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o2));
      }
      if (lane == 0) {
        s.thr[31] = do_ground ? mn : -__int_as_float(0x7f800000);
        s.thr[30] = do_ground ? mx : -__int_as_float(0x7f800000);
      }
    }
    __syncthreads();
    u32 C = 0;
    {
      constexpr u32 CH = 4u * CMAX;                       // rows per chunk
      constexpr u32 RPT = CH / kFrameThreads;             // rows (mask words) per thread: 8, 16 or 32
      static_assert(CH % kFrameThreads == 0 && RPT % 8 == 0 && RPT <= 32, "chunk rows per thread");
      static_assert(sizeof(s.u) >= CH * sizeof(unsigned short), "live-row list must fit the scratch union");
      u32* mword = reinterpret_cast<u32*>(&s.px[0]);      // [CH] aliases px, py, pz, pw (contiguous, 16 * CMAX bytes)
      unsigned short* liverow = reinterpret_cast<unsigned short*>(&s.u);   // [CH] chunk-local indices of the live rows
      const u32 nrows = (npts + 31u) >> 5;
      const float thr_min = s.thr[31], thr_max = s.thr[30];
      const u32 key_min = f2ord(thr_min);
      const bool skipping = do_ground && a.rowmax != nullptr;
      u32 mnx = 0xFFFFFFFFu, mny = 0xFFFFFFFFu, mnz = 0xFFFFFFFFu, mxx = 0, mxy = 0, mxz = 0;
      u32 gkept = 0;
      for (u32 cb = 0; cb < nrows; cb += CH) {
        const u32 nr = nrows - cb < CH ? nrows - cb : CH;
        const u32 r0 = tid * RPT;                          // this thread's first row of the chunk
        u32 w[RPT];
        if (!fused) {
          // keep bits from keep_mask_kernel
#pragma unroll
          for (u32 q = 0; q < RPT; q += 4) {
            uint4 m = make_uint4(0u, 0u, 0u, 0u);
            if (r0 + q < nr) m = *reinterpret_cast<const uint4*>(a.mask + (u64)tile0 * kTileWords + cb + r0 + q);
            w[q] = m.x; w[q + 1] = m.y; w[q + 2] = m.z; w[q + 3] = m.w;
          }
        } else {
          // (a) which rows can hold a survivor: a row whose highest z lies below the lowest ground threshold of
          // any sector is ground in every sector (z < thr_min <= thr[s], exact) and is never loaded
          u32 live = 0;
#pragma unroll
          for (u32 q = 0; q < RPT; q += 4) {
            uint4 m = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            if (skipping && r0 + q < nr)
              m = __ldcg(reinterpret_cast<const uint4*>(a.rowmax + (u64)tile0 * kTileWords + cb + r0 + q));
            live |= (m.x >= key_min ? 1u : 0u) << q;
            live |= (m.y >= key_min ? 1u : 0u) << (q + 1);
            live |= (m.z >= key_min ? 1u : 0u) << (q + 2);
            live |= (m.w >= key_min ? 1u : 0u) << (q + 3);
          }
          if (r0 + RPT > nr) live &= r0 < nr ? (0xFFFFFFFFu >> (32u - (nr - r0))) : 0u;
#pragma unroll
          for (u32 q = 0; q < RPT; ++q) mword[r0 + q] = 0u;
          u32 nlive;
          u32 lp = block_excl_scan<T>(__popc(live), s.wsum, nlive);
          while (live) {
            const u32 b = (u32)__ffs(live) - 1u;
            live &= live - 1;
            liverow[lp++] = (unsigned short)(r0 + b);
          }
          if (tid == 0) s.nlive = nlive;
          __syncthreads();
          CP_PHASE("live rows");
          // (b) evaluate the live rows, KB per warp and round: all loads of a round are in flight together
          constexpr u32 KB = T == 256 ? 4 : 8;   // (the 256-thread variant runs 4 CTAs per SM: 64 registers)
          for (u32 e0 = (u32)warp * KB; e0 < nlive; e0 += (kFrameThreads / 32) * KB) {
            float4 p[KB];
            u32 row[KB];
#pragma unroll
            for (u32 j = 0; j < KB; ++j) {
              row[j] = e0 + j < nlive ? (u32)liverow[e0 + j] : 0xFFFFFFFFu;
              const u32 i = (cb + row[j]) * 32u + lane;
              p[j] = (row[j] != 0xFFFFFFFFu && i < npts) ? load_point<MODE>(a.in, first + i, a.layout)
                                                          : make_float4(0.f, 0.f, -__int_as_float(0x7f800000), 0.f);
            }
#pragma unroll
            for (u32 j = 0; j < KB; ++j) {
              if (row[j] == 0xFFFFFFFFu) break;
              const u32 i = (cb + row[j]) * 32u + lane;
              const u32 bal = __ballot_sync(kFull, keep_point(p[j], i < npts, a.crop, a.gk, s.thr, thr_min, thr_max, gkept));
              if (lane == 0) mword[row[j]] = bal;
            }
          }
          __syncthreads();
          CP_PHASE("pass 2");
          if (tid == 0 && nlive) atomicAdd(a.rows_loaded, nlive);
#pragma unroll
          for (u32 q = 0; q < RPT; ++q) w[q] = mword[r0 + q];
        }
        // (c) ordered survivor indices of the chunk (k0 is free until the sort)
        u32 cnt = 0;
#pragma unroll
        for (u32 q = 0; q < RPT; ++q) cnt += __popc(w[q]);
        u32 total;
        u32 pos = C + block_excl_scan<T>(cnt, s.wsum, total);
        if (cnt && C + total <= (u32)CMAX) {
#pragma unroll
          for (u32 q = 0; q < RPT; ++q) {
            u32 bits = w[q];
            while (bits) {
              const u32 b = (u32)__ffs(bits) - 1u;
              bits &= bits - 1;
              s.k0[pos++] = (cb + r0 + q) * 32u + b;
            }
          }
        }
        C += total;
        __syncthreads();   // mword / liverow are rewritten by the next chunk
      }
      CP_PHASE("expand");
      if (do_ground) {
        gkept = __reduce_add_sync(kFull, gkept);
        if (lane == 0 && gkept) atomicAdd(&s.gkept, gkept);
      }
      __syncthreads();
      if (do_ground && tid == 0) a.gcount[f] = s.gkept;
      // all threads fetch the survivors in parallel (independent loads, no serial chains; L2 hits after (b))
      if (C <= (u32)CMAX) {
#pragma unroll 4
        for (u32 i = tid; i < C; i += kFrameThreads) {
          const float4 p = load_point<MODE>(a.in, first + s.k0[i], a.layout);
          s.px[i] = p.x; s.py[i] = p.y; s.pz[i] = p.z; s.pw[i] = p.w;
          const u32 kx = f2ord(p.x), ky = f2ord(p.y), kz = f2ord(p.z);
          mnx = min(mnx, kx); mxx = max(mxx, kx);
          mny = min(mny, ky); mxy = max(mxy, ky);
          mnz = min(mnz, kz); mxz = max(mxz, kz);
        }
      }
      (void)ntiles;
      mnx = __reduce_min_sync(kFull, mnx); mny = __reduce_min_sync(kFull, mny);
      mnz = __reduce_min_sync(kFull, mnz); mxx = __reduce_max_sync(kFull, mxx);
      mxy = __reduce_max_sync(kFull, mxy); mxz = __reduce_max_sync(kFull, mxz);
      if (lane == 0 && mxx >= mnx) {
        atomicMin(&s.bbox[0], mnx); atomicMin(&s.bbox[1], mny); atomicMin(&s.bbox[2], mnz);
        atomicMax(&s.bbox[4], mxx); atomicMax(&s.bbox[5], mxy); atomicMax(&s.bbox[6], mxz);
      }
    }
    // the zero points the ground node pads with (src/ground_removal.cpp:79) as ONE record
    // with multiplicity, appended after the real survivors
    const u32 pad_idx = a.pad_survives ? C : 0xFFFFFFFFu;
    if (a.pad_survives) {
      if (tid == 0 && C < (u32)CMAX) {
        s.px[C] = s.py[C] = s.pz[C] = s.pw[C] = 0.0f;
        const u32 kz = f2ord(0.0f);
        atomicMin(&s.bbox[0], kz); atomicMin(&s.bbox[1], kz); atomicMin(&s.bbox[2], kz);
        atomicMax(&s.bbox[4], kz); atomicMax(&s.bbox[5], kz); atomicMax(&s.bbox[6], kz);
      }
      ++C;
    }
    __syncthreads();

    CP_PHASE("fetch+bbox");
    // ---- S0b: VoxelGrid setup from the survivors' bounding box (thread 0)
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
          const float mn = ord2f(s.bbox[k]), mx = ord2f(s.bbox[4 + k]);
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
      a.ncrop_f[f] = C;
      atomicMax(&a.ctl->fast_max_c, C);
    }
    __syncthreads();
    u32 V = 0;
    const bool slow0 = s.slow != 0;

    if (!slow0) {
      // ---- S1: voxel keys (PCL idx) with the survivor position as the value
      const VoxelFrame vfr = s.vfr;
      for (u32 i = tid; i < C; i += kFrameThreads) {
        const i32 i0 = (i32)__fsub_rn(floorf(__fmul_rn(s.px[i], a.vk.inv[0])), (float)vfr.min_b[0]);
        const i32 i1 = (i32)__fsub_rn(floorf(__fmul_rn(s.py[i], a.vk.inv[1])), (float)vfr.min_b[1]);
        const i32 i2 = (i32)__fsub_rn(floorf(__fmul_rn(s.pz[i], a.vk.inv[2])), (float)vfr.min_b[2]);
        s.k0[i] = (u32)i0 + (u32)i1 * vfr.mul1 + (u32)i2 * vfr.mul2;
        s.v0[i] = (unsigned short)i;
      }
      CP_PHASE("keys");
      // ---- S2: stable LSD radix sort in shared memory, 8-bit digits, only the live key bits.
      // Warp w owns a contiguous chunk and walks it 32 items at a time, so ranks keep the
      // input (= point) order inside a voxel.
      constexpr int kWarps = kFrameThreads / 32;
      constexpr int kMaxRounds = CMAX / kFrameThreads;
      const u32 passes = (vfr.bits + 7u) >> 3;
      const u32 chunk = ((C + kFrameThreads - 1) / kFrameThreads) * 32;
      for (u32 pass = 0; pass < passes; ++pass) {
        const u32* ksrc = (pass & 1) ? s.k1 : s.k0;
        u32* kdst = (pass & 1) ? s.k0 : s.k1;
        const unsigned short* vsrc = (pass & 1) ? s.v1 : s.v0;
        unsigned short* vdst = (pass & 1) ? s.v0 : s.v1;
        const u32 shift = pass * 8;
        for (u32 q = tid; q < kWarps * 256; q += kFrameThreads) (&s.u.sort.whist[0][0])[q] = 0;
        __syncthreads();
        u32 key[kMaxRounds], off[kMaxRounds], vmask = 0;
        unsigned short val[kMaxRounds];
#pragma unroll
        for (int r = 0; r < kMaxRounds; ++r) {
          const u32 i = warp * chunk + r * 32 + lane;
          const bool valid = ((u32)r * 32 < chunk) && i < C;
          key[r] = valid ? ksrc[i] : 0u;
          val[r] = valid ? vsrc[i] : (unsigned short)0;
          const u32 d = valid ? ((key[r] >> shift) & 0xFFu) : 0x100u;
          const u32 peers = __match_any_sync(kFull, d);
          const int leader = __ffs(peers) - 1;
          u32 old = 0;
          if (valid && lane == leader) {
            old = s.u.sort.whist[warp][d];
            s.u.sort.whist[warp][d] = old + __popc(peers);
          }
          old = __shfl_sync(kFull, old, leader);
          off[r] = old + __popc(peers & lanemask_lt());
          vmask |= valid ? (1u << r) : 0u;
          __syncwarp();
        }
        __syncthreads();
        u32 run = 0;
        if (tid < 256) {
#pragma unroll
          for (int w = 0; w < kWarps; ++w) {
            const u32 t = s.u.sort.whist[w][tid];
            s.u.sort.whist[w][tid] = run;
            run += t;
          }
        }
        u32 tot;
        const u32 gb = block_excl_scan<T>(run, s.wsum, tot);
        if (tid < 256) s.u.sort.gbase[tid] = gb;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kMaxRounds; ++r) {
          if (vmask & (1u << r)) {
            const u32 d = (key[r] >> shift) & 0xFFu;
            const u32 pos = s.u.sort.gbase[d] + s.u.sort.whist[warp][d] + off[r];
            kdst[pos] = key[r];
            vdst[pos] = val[r];
          }
        }
        __syncthreads();
      }
      CP_PHASE("sort");
      const u32* ks = (passes & 1) ? s.k1 : s.k0;
      // ---- S3: segment heads -> voxel ids
      const u32 per = (C + kFrameThreads - 1) / kFrameThreads;
      const u32 r0 = tid * per, r1 = min(r0 + per, C);
      u32 cnt = 0;
      for (u32 r = r0; r < r1; ++r) cnt += (r == 0 || ks[r] != ks[r - 1]) ? 1u : 0u;
      u32 total;
      u32 vid = block_excl_scan<T>(cnt, s.wsum, total);
      V = total;
      if (V <= (u32)VMAX) {
        for (u32 r = r0; r < r1; ++r)
          if (r == 0 || ks[r] != ks[r - 1]) s.vstart[vid++] = r;
        if (tid == 0) s.vstart[V] = C;
      } else if (tid == 0) {
        s.slow = 1;
      }
      if (tid == 0) atomicMax(&a.ctl->fast_max_v, V);
      __syncthreads();
    }
    const bool slow = s.slow != 0;
    if (slow) V = 0;

    CP_PHASE("heads");
    // ---- voxel offsets across frames: only the parity taps need them (look-back in frame
    // order); the production path keeps frames independent
    if (a.tap_vox) {
      if (warp == 0) {
        const u32 e = lookback_exclusive(a.desc_v, f, V);
        if (lane == 0) {
          s.v_excl = e;
          a.v_off[f] = e;
          if (f + 1 == a.n_frames) {
            a.v_off[f + 1] = e + V;
            a.ctl->n_vox = e + V;
          }
        }
      }
    } else if (tid == 0) {
      s.v_excl = 0;
    }
    if (tid == 0) {
      a.nvox_f[f] = V;
      if (slow) atomicAdd(&a.ctl->fast_overflow, 1u);
    }
    __syncthreads();
    const u32 v_excl = s.v_excl;

    u32 K = 0;
    if (!slow && V > 0) {
      // ---- S4: voxel centroids — sequential fp32 sums in record (= point) order
      const u32 passes4 = (s.vfr.bits + 7u) >> 3;
      const u32* ks = (passes4 & 1) ? s.k1 : s.k0;
      const unsigned short* vs = (passes4 & 1) ? s.v1 : s.v0;
      // NOTE: the voxel arrays alias the sort scratch; the sort is complete (barrier above)
      float mxv = 0.f, myv = 0.f, mzv = 0.f;
      const u32 gcount_f = a.pad_survives ? (fused ? s.gkept : a.gcount[f]) : 0u;
      for (u32 v = tid; v < V; v += kFrameThreads) {
        const u32 b = s.vstart[v], e = s.vstart[v + 1];
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        u32 cnt = e - b;
        for (u32 r = b; r < e; ++r) {
          const u32 i = vs[r];
          sx = __fadd_rn(sx, s.px[i]);
          sy = __fadd_rn(sy, s.py[i]);
          sz = __fadd_rn(sz, s.pz[i]);
          si = __fadd_rn(si, s.pw[i]);
          if (i == pad_idx) cnt += (npts - gcount_f) - 1u;
        }
        const float c = (float)cnt;
        mxv = __fdiv_rn(sx, c); myv = __fdiv_rn(sy, c); mzv = __fdiv_rn(sz, c);
        s.u.vox.vx[v] = mxv; s.u.vox.vy[v] = myv; s.u.vox.vz[v] = mzv;
        s.u.vox.parent[v] = v;
        s.u.vox.csize[v] = 0;
        if (a.tap_vox && v_excl + v < a.tap_vox_cap)
          a.tap_vox[v_excl + v] = make_float4(mxv, myv, mzv, __fdiv_rn(si, c));
      }
      if (a.tap_keys) {
        for (u32 r = tid; r < C; r += kFrameThreads) {
          a.tap_keys[c0 + r] = ks[r];
          a.tap_order[c0 + r] = c0 + vs[r];
        }
      }
      CP_PHASE("means");
      // ---- S5: connected components of "L2_Simple(i,j) < r2".
      // Sweep: voxels are binned along the longer horizontal axis of the bounding box into
      // cells of edge >= 1.01 * tolerance (counting sort), so a voxel only meets the voxels of
      // its own and the previous cell.  A warp resolves a row cooperatively: roots of all hits,
      // one warp-wide minimum, and a CAS only for roots that still differ.
      if (tid == 0) {
        const float ex = __fsub_rn(ord2f(s.bbox[4]), ord2f(s.bbox[0]));
        const float ey = __fsub_rn(ord2f(s.bbox[5]), ord2f(s.bbox[1]));
        const u32 axis = ey > ex ? 1u : 0u;
        const float ext = axis ? ey : ex;
        float inv = a.ck.inv_h;
        u32 ncell = (u32)(ext * inv) + 2u;
        if (ncell > (u32)VMAX) {             // very long scene: widen the cells, still >= tolerance
          inv = (float)(VMAX - 2) / ext;
          ncell = (u32)VMAX;
        }
        s.sw_axis = axis;
        s.sw_ncell = ncell;
        s.sw_min = ord2f(s.bbox[axis]);
        s.sw_inv = inv;
      }
      __syncthreads();
      const u32 ncell = s.sw_ncell;
      for (u32 c = tid; c <= ncell; c += kFrameThreads) s.vstart[c] = 0;
      for (u32 c = tid; c < ncell; c += kFrameThreads) s.u.vox.label[c] = 0;
      __syncthreads();
      for (u32 v = tid; v < V; v += kFrameThreads) {
        const float coord = s.sw_axis ? s.u.vox.vy[v] : s.u.vox.vx[v];
        i32 c = (i32)floorf((coord - s.sw_min) * s.sw_inv);
        c = c < 0 ? 0 : (c >= (i32)ncell ? (i32)ncell - 1 : c);
        s.u.vox.vcell[v] = (unsigned short)c;
        atomicAdd(&s.vstart[c], 1u);
      }
      __syncthreads();
      {
        const u32 perc = (ncell + kFrameThreads - 1) / kFrameThreads;
        const u32 q0 = tid * perc, q1 = min(q0 + perc, ncell);
        u32 sum = 0;
        for (u32 q = q0; q < q1; ++q) sum += s.vstart[q];
        u32 tot;
        u32 run = block_excl_scan<T>(sum, s.wsum, tot);
        for (u32 q = q0; q < q1; ++q) {
          const u32 t = s.vstart[q];
          s.vstart[q] = run;
          run += t;
        }
        if (tid == 0) s.vstart[ncell] = V;
      }
      __syncthreads();
      for (u32 v = tid; v < V; v += kFrameThreads) {
        const u32 c = s.u.vox.vcell[v];
        s.u.vox.perm[s.vstart[c] + atomicAdd(&s.u.vox.label[c], 1u)] = (unsigned short)v;
      }
      __syncthreads();
      CP_PHASE("binning");
#ifdef CP_UNION_LEVEL
      __shared__ u32 dbg_cnt[2];
      if (tid < 2) dbg_cnt[tid] = 0;
      __syncthreads();
#endif
      // eight lanes per row, candidates strided over them: a few dozen rows run side by side per
      // warp and a crowded cell (many candidates) does not stall the warp the way one thread per row
      // does (measured: rows average 26 candidates but the longest have > 100).  The lanes of a row
      // need no communication: each carries the row's root (or an ancestor of it) in a register.
      constexpr u32 kLanesPerRow = 8;
      u32 n_visited = 0, n_tested = 0;   // statistics (cp_last_pairs): registers, one atomic per warp at the end
      for (u32 r = tid / kLanesPerRow; r < V; r += kFrameThreads / kLanesPerRow) {
        const u32 i = s.u.vox.perm[r];
        const u32 ci = s.u.vox.vcell[i];
        const u32 lo = s.vstart[ci ? ci - 1 : 0];
        const float xi = s.u.vox.vx[i], yi = s.u.vox.vy[i], zi = s.u.vox.vz[i];
        u32 ri = smem_find(s.u.vox.parent, i);
#ifdef CP_UNION_LEVEL
        u32 dbg_hits = 0, dbg_cand = (tid % kLanesPerRow == 0) ? r - lo : 0;
#endif
        for (u32 jp = lo + tid % kLanesPerRow; jp < r; jp += kLanesPerRow) {
          const u32 j = s.u.vox.perm[jp];
          ++n_visited;
#ifndef CP_UNION_LEVEL
          // j already hangs under i's root: the edge cannot change anything, so neither the distance nor a
          // find is needed.  The voxels of one cone are mutual neighbours (a clique of up to ~100 at close
          // range); after its first rows almost every candidate takes this one-load exit.
          if (((volatile u32*)s.u.vox.parent)[j] == ri) continue;
#endif
          ++n_tested;
          if (l2_simple(xi, yi, zi, s.u.vox.vx[j], s.u.vox.vy[j], s.u.vox.vz[j]) < a.ck.r2) {
#if defined(CP_UNION_LEVEL) && CP_UNION_LEVEL == 0
            dbg_hits++;
#elif defined(CP_UNION_LEVEL) && CP_UNION_LEVEL == 1
            dbg_hits += smem_find(s.u.vox.parent, j);
#else
            // link the larger root under the smaller with one CAS on what we believe are the two roots; only
            // when the larger one has been linked meanwhile do we climb again (from where the CAS pointed us).
            // Linking under a node that is no longer a root is still sound: parents always have smaller
            // indices, so there are no cycles and every tree's root is the smallest index of its component.
            u32 rj = smem_find(s.u.vox.parent, j);
            if (rj != j) ((volatile u32*)s.u.vox.parent)[j] = rj;  // j is not a root: point it at its root
            while (rj != ri) {
              const u32 hi = ri > rj ? ri : rj, sm = ri > rj ? rj : ri;
              const u32 old = atomicCAS(&s.u.vox.parent[hi], hi, sm);
              if (old == hi) {
                ri = sm;
                break;
              }
              if (hi == ri) ri = smem_find(s.u.vox.parent, old);
              else rj = smem_find(s.u.vox.parent, old);
            }
#endif
          }
        }
#ifndef CP_UNION_LEVEL
        // ri is i itself or one of its ancestors: later rows that meet i take the one-load exit above
        if (ri != i) ((volatile u32*)s.u.vox.parent)[i] = ri;
#endif
#ifdef CP_UNION_LEVEL
        if (f == 0) {
          atomicAdd(&dbg_cnt[0], dbg_cand);
          atomicAdd(&dbg_cnt[1], dbg_hits);
        }
#endif
      }
      n_visited = __reduce_add_sync(kFull, n_visited);
      n_tested = __reduce_add_sync(kFull, n_tested);
      if (lane == 0 && n_visited) {
        atomicAdd(&a.ctl->pairs_visited, (unsigned long long)n_visited);
        atomicAdd(&a.ctl->pairs_tested, (unsigned long long)n_tested);
      }
      __syncthreads();
      CP_PHASE("union");
#ifdef CP_UNION_LEVEL
      if (tid == 0 && f == 0) printf("V=%u candidates=%u hits(or sum)=%u\n", V, dbg_cnt[0], dbg_cnt[1]);
#endif
      // ---- S6: labels (root = min voxel index of the component) and component sizes
      for (u32 v = tid; v < V; v += kFrameThreads) s.u.vox.label[v] = smem_find(s.u.vox.parent, v);
      __syncthreads();
      for (u32 v = tid; v < V; v += kFrameThreads) {
        atomicAdd(&s.u.vox.csize[s.u.vox.label[v]], 1u);
        if (a.tap_labels) a.tap_labels[v_excl + v] = (i32)s.u.vox.label[v];
      }
      __syncthreads();
      // ---- S7: kept roots in ascending order (ordered compaction over v)
      const u32 perv = (V + kFrameThreads - 1) / kFrameThreads;
      const u32 v0 = tid * perv, v1 = min(v0 + perv, V);
      u32 kc = 0, cc = 0;
      for (u32 v = v0; v < v1; ++v) {
        if (s.u.vox.label[v] == v) {
          ++cc;
          const u32 sz = s.u.vox.csize[v];
          if (sz >= a.ck.min_size && sz <= a.ck.max_size) ++kc;
        }
      }
      u32 ktotal, ctotal;
      u32 kpos = block_excl_scan<T>(kc, s.wsum, ktotal);
      block_excl_scan<T>(cc, s.wsum, ctotal);
      K = ktotal;
      for (u32 v = v0; v < v1; ++v) {
        if (s.u.vox.label[v] == v) {
          const u32 sz = s.u.vox.csize[v];
          if (sz >= a.ck.min_size && sz <= a.ck.max_size) s.vstart[kpos++] = v;  // vstart is free now
        }
      }
      if (tid == 0) s.n_comp = ctotal;
      __syncthreads();
    } else if (tid == 0) {
      s.n_comp = 0;
    }

    CP_PHASE("labels+kept");
    // ---- per-frame results go to the frame's own slot; pack_clusters_kernel compacts them
    if (tid == 0) {
      a.kcount_f[f] = K;
      a.ncomp_f[f] = s.n_comp;
      u32* fc = a.fc + (u64)f * 8;   // cp_frame_counters
      fc[0] = npts;
      fc[1] = a.counted_ground ? (fused ? s.gkept : a.gcount[f]) : npts;   // G (= N without ground removal)
      fc[2] = C;
      fc[3] = V;
      fc[4] = s.n_comp;
      fc[5] = K;
      fc[6] = s.vfr.bits;
      fc[7] = s.vfr.passthrough;
      if (a.direct_out) {            // what pack_clusters_kernel would write for a one-frame batch
        a.k_off[0] = 0;
        a.k_off[1] = K;
        a.ctl->n_clusters = K;
        if (K > a.direct_cap) atomicOr(&a.ctl->error, kErrVoxels);
      }
    }
    // ---- S8/S9: centroid (src/cone_detection.cpp:261-273) and canonical rank, warp per cluster.
    // Lanes find the members 32 voxels at a time; the fp32 sums stay sequential in ascending
    // voxel index (every lane carries the same running sum).
    if (K > 0) {
      ClusterRec* slot = a.direct_out ? a.direct_out : a.slots + (u64)f * VMAX;
      const u32 slot_cap = a.direct_out ? a.direct_cap : (u32)VMAX;
      for (u32 k = warp; k < K; k += kFrameThreads / 32) {
        const u32 root = s.vstart[k];
        const u32 size = s.u.vox.csize[root];
        float x = 0.0f, y = 0.0f;
        u32 seen = 0;
        for (u32 base = root; base < V && seen < size; base += 32) {
          const u32 v = base + lane;
          u32 members = __ballot_sync(kFull, v < V && s.u.vox.label[v] == root);
          seen += __popc(members);
          while (members) {
            const u32 m = base + (u32)__ffs(members) - 1u;
            members &= members - 1;
            x = __fadd_rn(x, s.u.vox.vx[m]);
            y = __fadd_rn(y, s.u.vox.vy[m]);
          }
        }
        u32 rank = 0;  // size descending, then min index ascending (kept roots are ascending)
        for (u32 q = lane; q < K; q += 32) {
          const u32 sq = s.u.vox.csize[s.vstart[q]];
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
          if (rank < slot_cap) slot[rank] = o;
        }
      }
    }
#ifdef CP_PHASE_CLOCKS
    __syncthreads();
    if (tid == 0 && f == 0) {
      long long prev = clk_0;
      for (int q = 0; q < clk_n; ++q) {
        printf("%-12s %8lld cyc\n", clk_name[q], clk_t[q] - prev);
        prev = clk_t[q];
      }
      printf("%-12s %8lld cyc  (total %lld)\n", "tail", clock64() - prev, clock64() - clk_0);
    }
#endif
  }
}

template <int CMAX, int VMAX, int MODE, int T>
__global__ void __launch_bounds__(T, T == 256 ? 4 : 1) frame_backend_kernel(FrameArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FrameSmem<CMAX, VMAX, T>& s = *reinterpret_cast<FrameSmem<CMAX, VMAX, T>*>(smem_raw);
  while (true) {   // frames are handed out by ticket
    __syncthreads();
    if (threadIdx.x == 0) s.frame = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const u32 f = s.frame;
    if (f >= a.n_frames) break;
    frame_process<CMAX, VMAX, MODE, T>(a, s, f);
  }
}

// compaction of the per-frame result slots into the packed cluster list + offsets.
// One CTA per frame; every CTA sums the counts of the frames before it (F is small).
__global__ void __launch_bounds__(256) pack_clusters_kernel(u32 n_frames, u32 slot_stride,
                                                            const u32* __restrict__ kcount_f,
                                                            const ClusterRec* __restrict__ slots,
                                                            u32* __restrict__ k_off, ClusterRec* __restrict__ out,
                                                            u32 out_cap, Ctl* ctl) {
  __shared__ u32 wsum[8];
  __shared__ u32 s_off;
  const u32 f = blockIdx.x;
  u32 part = 0;
  for (u32 i = threadIdx.x; i < f; i += blockDim.x) part += kcount_f[i];
  part = __reduce_add_sync(kFull, part);
  if (lane_id() == 0) wsum[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 t = 0;
    for (int w = 0; w < 8; ++w) t += wsum[w];
    s_off = t;
    k_off[f] = t;
    if (f + 1 == n_frames) {
      const u32 total = t + kcount_f[f];
      k_off[n_frames] = total;
      ctl->n_clusters = total;
      if (total > out_cap) atomicOr(&ctl->error, kErrVoxels);
    }
  }
  __syncthreads();
  const u32 off = s_off, k = kcount_f[f];
  for (u32 i = threadIdx.x; i < k; i += blockDim.x)
    if (off + i < out_cap) out[off + i] = slots[(u64)f * slot_stride + i];
}

}  // namespace cp
