// fps.cu -- furthest point sampling as a thread-block-cluster kernel (sm_100a).
//
// Replaces furthest_point_sampling_kernel (reference sampling_gpu.cu:69-229), which runs ONE
// 512-thread block per scene and read-modify-writes a global `temp` array every round.
//
// Design (B200-first):
//  * one CLUSTER of C CTAs per scene (C = 1..16); every point's xyz and running min-distance
//    stay in REGISTERS for the whole kernel (P points per thread) -- no global traffic at all
//    inside the npoint-1 sequential rounds;
//  * per round: P fused distance/min updates per thread, a 2-instruction warp arg-max
//    (redux.sync max on the distance bits, redux.sync min on the tie-break key), one
//    __syncthreads for the CTA-level candidate table, then every CTA pushes its candidate
//    (distance, index, x, y, z) into every peer's shared memory with st.async, whose
//    completion is counted on the peer's mbarrier (no cluster barrier, no fence) -- each CTA
//    then reduces the C candidates redundantly, so the winner's coordinates are already
//    on-chip for the next round;
//  * bit-exactness with the reference's block-tree arg-max: among equal maxima the reference
//    keeps the candidate with the smallest (bit-reversed thread id, k div T), T =
//    opt_n_threads(N).  Points are assigned to threads in exactly that order
//    ("virtual index" v = bitrev(k mod T) * Q + k div T, Q = ceil(N/T)), so "max distance, then
//    min v" reproduces the tree, and each thread's strict '>' scan reproduces the per-thread rule;
//  * the |p|^2 <= 1e-3 skip (sampling_gpu.cu:100-101, compared in double) is evaluated once at
//    load time: skipped / out-of-range slots get min-distance -1 and can never win.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace spc {

constexpr int kMaxCluster = 16;

struct FpsParams {
  const float *xyz;   // (B,N,3)
  int32_t *idx;       // (B,npoint)
  float *new_xyz;     // (B,npoint,3) or nullptr
  int N, npoint;
  int T, log2T, Q;    // reference thread count, its log2, ceil(N/T)
  const int *ordered_ok;  // per scene: 1 = "FPS(xyz)[0:npoint] is provably 0..npoint-1" (or nullptr)
  int *strict_out;        // per scene (or nullptr): 1 = every pick was the strict unique maximum (culled kernels)
};

__device__ __forceinline__ int fps_v_to_k(unsigned v, int Q, int T, int log2T) {
  const unsigned r = v / (unsigned)Q;
  const unsigned q = v - r * (unsigned)Q;
  const unsigned res = log2T ? (__brev(r) >> (32 - log2T)) : 0u;
  return (int)(q * (unsigned)T + res);
}

// ---- PTX helpers: cluster address mapping, mbarrier, async DSMEM store --------------------------
__device__ __forceinline__ uint32_t fps_s2u(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void fps_mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fps_s2u(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fps_mbar_arm(uint64_t *bar, unsigned tx_bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fps_s2u(bar)), "r"(tx_bytes)
               : "memory");
}
// NB: plain (cta-scope) try_wait.  A cluster-scope acquire here compiles to an L1 invalidate
// (CCTL.IVALL) on every poll; completion of a transaction barrier already makes the st.async
// payload visible (same contract as a TMA load).
__device__ __forceinline__ void fps_mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "FPS_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra FPS_DONE;\n"
      "bra FPS_WAIT;\n"
      "FPS_DONE:\n"
      "}\n" ::"r"(fps_s2u(bar)),
      "r"(parity)
      : "memory");
}
// 16-byte + 4-byte remote stores into a peer CTA's shared memory; each store signals its bytes
// on the peer's mbarrier (complete_tx), so the receiver needs no fence and no cluster barrier
// (cg::cluster_group::sync() costs a MEMBAR.ALL.GPU + L1 invalidate per round: measured 40 % of
// the kernel in the first version).
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint32_t a, uint32_t b, uint32_t c,
                                            uint32_t d, uint32_t remote_bar) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::
          "r"(remote_addr),
      "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar)
      : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t remote_addr, uint32_t a, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
               "r"(a), "r"(remote_bar)
               : "memory");
}

// candidate record exchanged between warps / CTAs: 32 bytes, 16-byte aligned
struct __align__(16) FpsCand {
  unsigned key;   // distance bits + 1, 0 = "no valid point"
  int k;          // point index
  float x, y;
  float z;
  unsigned pad[3];
};

// XYZ_REGS: coordinates in registers (fast path); otherwise they are read from shared memory
// every round (large-N fallback, P up to 32 with 512 threads).
//
// Tie-break without a second reduction: points are laid out so that the virtual index v grows
// with (cluster rank, warp, lane, slot).  "Largest distance, then smallest v" therefore is
// "largest key, then FIRST in (rank, warp, lane, slot) order" -- one redux.max + one ballot/ffs
// per level, and the per-thread strict '>' keeps the lowest slot.
template <int P, int THREADS, bool XYZ_REGS>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 ? 2 : 1)) fps_cluster_kernel(const FpsParams p) {
  constexpr int NW = THREADS / 32;
  extern __shared__ float s_xyz[];  // [3][P][THREADS] SoA copy of this CTA's points + [P][THREADS] index
  float *sx = s_xyz, *sy = s_xyz + P * THREADS, *sz = s_xyz + 2 * P * THREADS;
  int *sk = reinterpret_cast<int *>(s_xyz + 3 * P * THREADS);
  __shared__ FpsCand w_cand[2][NW];            // per-warp candidates (double buffered)
  __shared__ FpsCand c_cand[2][kMaxCluster];   // per-CTA candidates, written by peers via DSMEM
  __shared__ __align__(8) uint64_t c_bar[2];   // "all C candidates of this buffer have landed"

  cg::cluster_group cluster = cg::this_cluster();
  const unsigned C = cluster.num_blocks();
  const unsigned rank = cluster.block_rank();
  const int scene = blockIdx.y;
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31u, warp = tid >> 5;
  const float *xyz = p.xyz + (size_t)scene * p.N * 3;
  const unsigned g = rank * THREADS + tid;      // thread id within the cluster
  const unsigned v0 = g * P;                    // first virtual index owned by this thread
  const unsigned V = (unsigned)p.T * (unsigned)p.Q;

  // Verified shortcut (see fps_prefix_check_kernel): the input is an FPS-ordered prefix, so the
  // npoint sequential rounds would return 0..npoint-1.  Every CTA of the cluster reads the same
  // flag, so they all leave together.
  if (p.ordered_ok != nullptr && p.ordered_ok[scene] != 0) {
    if (g == 0 && p.strict_out) p.strict_out[scene] = 1;   // proven: every pick was a strict unique maximum
    int32_t *oidx = p.idx + (size_t)scene * p.npoint;
    for (unsigned j = g; j < (unsigned)p.npoint; j += C * THREADS) oidx[j] = (int)j;
    if (p.new_xyz) {
      float *o = p.new_xyz + (size_t)scene * p.npoint * 3;
      for (unsigned e = g; e < 3u * (unsigned)p.npoint; e += C * THREADS) o[e] = __ldg(xyz + e);
    }
    return;
  }

  float x[P], y[P], z[P], t[P];
#pragma unroll
  for (int s = 0; s < P; ++s) {
    const unsigned v = v0 + s;
    float px = 0.f, py = 0.f, pz = 0.f, pt = -1.0f;
    int k = 0;
    if (v < V) {
      k = fps_v_to_k(v, p.Q, p.T, p.log2T);
      if (k < p.N) {
        px = __ldg(xyz + 3 * k + 0);
        py = __ldg(xyz + 3 * k + 1);
        pz = __ldg(xyz + 3 * k + 2);
        const float mag = __fmaf_rn(pz, pz, __fmaf_rn(px, px, __fmul_rn(py, py)));
        pt = ((double)mag <= 1e-3) ? -1.0f : 1e10f;   // reference compares in double (F5)
      }
    }
    sx[s * THREADS + tid] = px;
    sy[s * THREADS + tid] = py;
    sz[s * THREADS + tid] = pz;
    sk[s * THREADS + tid] = k;
    if (XYZ_REGS) { x[s] = px; y[s] = py; z[s] = pz; }
    t[s] = pt;
  }
  // point 0 is always the first pick (sampling_gpu.cu:85-86) and the fallback when no valid
  // point exists (every thread reports best=-1, besti=0 => dists_i[0] == 0).
  const float p0x = __ldg(xyz + 0), p0y = __ldg(xyz + 1), p0z = __ldg(xyz + 2);
  float ox = p0x, oy = p0y, oz = p0z;
  int32_t *idx = p.idx + (size_t)scene * p.npoint;
  float *nxyz = p.new_xyz ? p.new_xyz + (size_t)scene * p.npoint * 3 : nullptr;
  // the output writer sits in the LAST warp so that the first warp (which drives the cluster
  // exchange) never lags behind
  const bool writer = (rank == 0 && tid == THREADS - 1);
  if (writer) {
    idx[0] = 0;
    if (nxyz) { nxyz[0] = ox; nxyz[1] = oy; nxyz[2] = oz; }
  }
  const unsigned tx_bytes = 20u * C;
  uint32_t r_slot0 = 0, r_slot1 = 0, r_bar0 = 0, r_bar1 = 0;   // my slot / barrier in peer `lane`
  if (C > 1) {
    if (tid == 0) {
      fps_mbar_init(&c_bar[0], 1);
      fps_mbar_init(&c_bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      fps_mbar_arm(&c_bar[0], tx_bytes);
      fps_mbar_arm(&c_bar[1], tx_bytes);
    }
    cluster.sync();   // peers' barriers are initialised before anyone stores into them
    if (warp == 0 && lane < C) {
      r_slot0 = mapa_cluster(fps_s2u(&c_cand[0][rank]), lane);
      r_slot1 = mapa_cluster(fps_s2u(&c_cand[1][rank]), lane);
      r_bar0 = mapa_cluster(fps_s2u(&c_bar[0]), lane);
      r_bar1 = mapa_cluster(fps_s2u(&c_bar[1]), lane);
    }
  } else {
    __syncthreads();
  }

  for (int j = 1; j < p.npoint; ++j) {
    const int buf = j & 1;
    float best = -1.0f;
    int bs = 0;
#pragma unroll
    for (int s = 0; s < P; ++s) {
      float d;
      if (XYZ_REGS) d = sqdist_ref(x[s], y[s], z[s], ox, oy, oz);
      else d = sqdist_ref(sx[s * THREADS + tid], sy[s * THREADS + tid], sz[s * THREADS + tid], ox, oy, oz);
      const float d2 = fminf(d, t[s]);
      t[s] = d2;
      if (d2 > best) { best = d2; bs = s; }   // strict '>': lowest slot (= lowest v) wins ties
    }
    // every lane fetches its own candidate (overlaps the reduction latency; no divergent chain)
    const float mx = sx[bs * THREADS + tid], my = sy[bs * THREADS + tid], mz = sz[bs * THREADS + tid];
    const int mk = sk[bs * THREADS + tid];
    // ---- warp arg-max: redux.max on the distance bits; first lane among equals wins ----------
    const unsigned key = best < 0.f ? 0u : __float_as_uint(best) + 1u;
    const unsigned wmax = __reduce_max_sync(0xffffffffu, key);
    // first lane among equals: redux.min on the lane id (26 cycles vs 66 for ballot + ffs)
    if (lane == __reduce_min_sync(0xffffffffu, key == wmax ? lane : 32u)) {
      FpsCand *e = &w_cand[buf][warp];
      *reinterpret_cast<uint4 *>(e) = make_uint4(wmax, (unsigned)mk, __float_as_uint(mx), __float_as_uint(my));
      e->z = mz;
    }
    __syncthreads();
    // ---- CTA arg-max, computed redundantly by every warp (first warp among equals wins) ------
    uint4 e4 = make_uint4(0u, 0u, 0u, 0u);
    float ez = 0.f;
    if (lane < NW) {
      e4 = *reinterpret_cast<const uint4 *>(&w_cand[buf][lane]);
      ez = w_cand[buf][lane].z;
    }
    unsigned bmax = __reduce_max_sync(0xffffffffu, e4.x);
    unsigned src = __reduce_min_sync(0xffffffffu, (lane < NW && e4.x == bmax) ? lane : 32u);
    unsigned wk = __shfl_sync(0xffffffffu, e4.y, src);
    unsigned wxb = __shfl_sync(0xffffffffu, e4.z, src);
    unsigned wyb = __shfl_sync(0xffffffffu, e4.w, src);
    float wz = __shfl_sync(0xffffffffu, ez, src);
    if (C > 1) {
      // ---- push this CTA's candidate to every CTA of the cluster; data + completion in one op --
      if (warp == 0 && lane < C) {
        const uint32_t rs = buf ? r_slot1 : r_slot0, rb = buf ? r_bar1 : r_bar0;
        st_async_v4(rs, bmax, wk, wxb, wyb, rb);
        st_async_b32(rs + 16, __float_as_uint(wz), rb);
      }
      fps_mbar_wait(&c_bar[buf], (unsigned)((j - 1) >> 1) & 1u);   // u-th use of this buffer
      if (tid == 0) fps_mbar_arm(&c_bar[buf], tx_bytes);           // re-arm for round j+2
      e4 = make_uint4(0u, 0u, 0u, 0u);
      ez = 0.f;
      if (lane < C) {
        e4 = *reinterpret_cast<const uint4 *>(&c_cand[buf][lane]);
        ez = c_cand[buf][lane].z;
      }
      bmax = __reduce_max_sync(0xffffffffu, e4.x);
      src = __reduce_min_sync(0xffffffffu, (lane < C && e4.x == bmax) ? lane : 32u);
      wk = __shfl_sync(0xffffffffu, e4.y, src);
      wxb = __shfl_sync(0xffffffffu, e4.z, src);
      wyb = __shfl_sync(0xffffffffu, e4.w, src);
      wz = __shfl_sync(0xffffffffu, ez, src);
    }
    int old = 0;
    if (bmax == 0u) { ox = p0x; oy = p0y; oz = p0z; }
    else { ox = __uint_as_float(wxb); oy = __uint_as_float(wyb); oz = wz; old = (int)wk; }
    if (writer) {
      idx[j] = old;
      if (nxyz) { nxyz[3 * j + 0] = ox; nxyz[3 * j + 1] = oy; nxyz[3 * j + 2] = oz; }
    }
  }
  if (writer && p.strict_out) p.strict_out[scene] = 0;   // this kernel does not track ties: "unknown"
  if (C > 1) cluster.sync();  // no CTA may exit while a peer can still store into its smem
}

// ================================================================================================
// Culled variant for large clouds.
//
// After the first few dozen picks a new centre only changes the min-distance of points within the
// current sampling radius, i.e. of a few per cent of the cloud -- yet the kernel above updates
// every point every round, and that distance math is 2/3 of its issue slots (which is what limits
// throughput once several scenes share an SM).  Here the points are first sorted along a Morton
// curve (fps_morton_sort_kernel), so that the 32*P points of a WARP are spatially compact; each
// warp keeps the bounding box of its points and its current best candidate.  If the new centre is
// farther from the box than the warp's largest min-distance, no min-distance in the warp can
// change (d >= box distance >= every t) and the warp just republishes its cached candidate.
// The test is conservative w.r.t. fp32 rounding (factor 1-1e-5), so the result is bit-identical.
// Because points are no longer laid out in the reference's tie-break order, ties are broken
// explicitly with the virtual index v (one extra redux per level); each thread's slots are sorted
// by v once at load time so that the strict '>' scan still keeps the lowest v.
// ================================================================================================
constexpr int FMS_THREADS = 1024;
constexpr int FMS_BINS = 32768;       // 32^3 Morton cells

__device__ __forceinline__ unsigned morton_spread5(unsigned x) {   // 5 bits -> every third bit
  x = (x | (x << 8)) & 0x0000100Fu;
  x = (x | (x << 4)) & 0x000010C3u;
  x = (x | (x << 2)) & 0x00001249u;   // bits 0,3,6,9,12
  return x;
}

// one CTA per scene: bounding box -> 15-bit Morton cell per point -> counting sort -> perm[b][N].
// N <= FMS_THREADS * FMS_ITEMS (the culled kernel's own capacity); every thread keeps the cell codes of
// its <= 40 points in registers (two per register) between the counting and the scatter pass, and
// loads are issued eight points at a time so that the three passes are not latency-bound.
constexpr int FMS_ITEMS = 40;
constexpr int FMS_BATCH = 8;

__global__ void __launch_bounds__(FMS_THREADS) fps_morton_sort_kernel(const float *__restrict__ xyz, int N,
                                                                      int32_t *__restrict__ perm) {
  extern __shared__ int s_cnt[];                 // [FMS_BINS]
  __shared__ float s_red[6][32];
  __shared__ float s_box[6];
  __shared__ int s_part[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *P = xyz + (size_t)b * N * 3;
  for (int c = tid; c < FMS_BINS; c += FMS_THREADS) s_cnt[c] = 0;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int i0 = 0; i0 < FMS_ITEMS; i0 += FMS_BATCH) {
    float v[FMS_BATCH][3];
#pragma unroll
    for (int i = 0; i < FMS_BATCH; ++i) {
      const int k = tid + (i0 + i) * FMS_THREADS;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[i][c] = k < N ? __ldg(P + 3 * k + c) : NAN;   // NaN: ignored by fmin/fmax
    }
#pragma unroll
    for (int i = 0; i < FMS_BATCH; ++i)
#pragma unroll
      for (int c = 0; c < 3; ++c) { mn[c] = fminf(mn[c], v[i][c]); mx[c] = fmaxf(mx[c], v[i][c]); }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
    if (lane == 0) { s_red[c][warp] = mn[c]; s_red[3 + c][warp] = mx[c]; }
  }
  __syncthreads();
  if (tid < 3) {
    float lo = s_red[tid][0], hi = s_red[3 + tid][0];
    for (int w = 1; w < FMS_THREADS / 32; ++w) { lo = fminf(lo, s_red[tid][w]); hi = fmaxf(hi, s_red[3 + tid][w]); }
    const float ext = hi - lo;
    s_box[tid] = lo;
    s_box[3 + tid] = (ext > 0.f && ext < INFINITY) ? 32.0f / ext : 0.f;
  }
  __syncthreads();
  const float bx = s_box[0], by = s_box[1], bz = s_box[2], gx = s_box[3], gy = s_box[4], gz = s_box[5];
  unsigned codes[FMS_ITEMS / 2];
#pragma unroll
  for (int i0 = 0; i0 < FMS_ITEMS; i0 += FMS_BATCH) {
    float v[FMS_BATCH][3];
#pragma unroll
    for (int i = 0; i < FMS_BATCH; ++i) {
      const int k = tid + (i0 + i) * FMS_THREADS;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[i][c] = k < N ? __ldg(P + 3 * k + c) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < FMS_BATCH; ++i) {
      const int k = tid + (i0 + i) * FMS_THREADS;
      const unsigned qx = (unsigned)min(31, max(0, (int)((v[i][0] - bx) * gx)));
      const unsigned qy = (unsigned)min(31, max(0, (int)((v[i][1] - by) * gy)));
      const unsigned qz = (unsigned)min(31, max(0, (int)((v[i][2] - bz) * gz)));
      const unsigned code = morton_spread5(qx) | (morton_spread5(qy) << 1) | (morton_spread5(qz) << 2);
      if (k < N) atomicAdd(&s_cnt[code], 1);
      if ((i & 1) == 0) codes[(i0 + i) >> 1] = code;
      else codes[(i0 + i) >> 1] |= code << 16;
    }
  }
  __syncthreads();
  constexpr int PER = FMS_BINS / FMS_THREADS;
  const int c0 = tid * PER;
  int sum = 0;
#pragma unroll 8
  for (int c = c0; c < c0 + PER; ++c) sum += s_cnt[c];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_part[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = s_part[lane], inc2 = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc2, o); if (lane >= o) inc2 += u; }
    s_part[lane] = inc2 - v;
  }
  __syncthreads();
  int run = s_part[warp] + incl - sum;
#pragma unroll 8
  for (int c = c0; c < c0 + PER; ++c) { const int n = s_cnt[c]; s_cnt[c] = run; run += n; }
  __syncthreads();
  int32_t *out = perm + (size_t)b * N;
#pragma unroll
  for (int i = 0; i < FMS_ITEMS; ++i) {
    const int k = tid + i * FMS_THREADS;
    const unsigned code = (codes[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
    if (k < N) out[atomicAdd(&s_cnt[code], 1)] = k;
  }
}

struct __align__(16) FpsCandV {
  unsigned key, v; int k; float x;
  float y, z; unsigned pad[2];
};

__device__ __forceinline__ void st_async_v2(uint32_t remote_addr, uint32_t a, uint32_t b, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::
                   "r"(remote_addr),
               "r"(a), "r"(b), "r"(remote_bar)
               : "memory");
}

// Several lanes hold the maximal key: the smallest virtual index wins.  Deliberately NOT inlined: ptxas
// otherwise if-converts the rare path into predicated redux that sit on every round's dependency chain.
__device__ __noinline__ unsigned fps_resolve_tie(bool hit, unsigned v, unsigned lane) {
  const unsigned vmin = __reduce_min_sync(0xffffffffu, hit ? v : 0xffffffffu);
  return __reduce_min_sync(0xffffffffu, (hit && v == vmin) ? lane : 32u);
}

// XYZ_SMEM = false: coordinates and min-distances in registers (128 regs, 100 KB smem: two CTAs per SM).
// XYZ_SMEM = true : only the min-distances stay in registers; coordinates are read from shared memory by
//   the (few) warps that are not culled, point indices are kept as uint16 and the tie-break index is
//   recomputed from them => <= 80 regs and 70 KB smem: THREE CTAs per SM, i.e. a scene occupies 2.7 SMs
//   instead of 4 for the duration of the call.
// TRACK = true additionally maintains p.strict_out (see below); compiled out otherwise (it costs registers).
template <int P, int THREADS, bool XYZ_SMEM, bool TRACK>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 ? (XYZ_SMEM ? 3 : 2) : 1))
fps_cull_kernel(const FpsParams p, const int32_t *__restrict__ perm_all) {
  constexpr int NW = THREADS / 32;
  extern __shared__ float s_xyz[];  // [3][P][THREADS] xyz + ([P][THREADS] k + [P][THREADS] v | [P][THREADS] k16)
  float *sx = s_xyz, *sy = s_xyz + P * THREADS, *sz = s_xyz + 2 * P * THREADS;
  // during the load-time sort the XYZ_SMEM variant borrows the (not yet written) x / y planes for v / k
  int *sk = XYZ_SMEM ? reinterpret_cast<int *>(sy) : reinterpret_cast<int *>(s_xyz + 3 * P * THREADS);
  unsigned *sv = XYZ_SMEM ? reinterpret_cast<unsigned *>(sx) : reinterpret_cast<unsigned *>(s_xyz + 4 * P * THREADS);
  uint16_t *sk16 = reinterpret_cast<uint16_t *>(s_xyz + 3 * P * THREADS);
  auto v_of_k = [&](unsigned k) -> unsigned {
    const unsigned res = p.log2T ? (__brev(k & (unsigned)(p.T - 1)) >> (32 - p.log2T)) : 0u;
    return res * (unsigned)p.Q + (k >> p.log2T);      // T is a power of two
  };
  __shared__ FpsCandV w_cand[2][NW];
  __shared__ FpsCandV c_cand[2][kMaxCluster];
  __shared__ __align__(8) uint64_t c_bar[2];

  cg::cluster_group cluster = cg::this_cluster();
  const unsigned C = cluster.num_blocks();
  const unsigned rank = cluster.block_rank();
  const int scene = blockIdx.y;
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31u, warp = tid >> 5;
  const float *xyz = p.xyz + (size_t)scene * p.N * 3;
  const int32_t *perm = perm_all + (size_t)scene * p.N;
  const unsigned g = rank * THREADS + tid;
  const unsigned q0 = g * P;                    // first Morton-sorted position owned by this thread
  if (p.ordered_ok != nullptr && p.ordered_ok[scene] != 0) {   // verified shortcut, as in fps_cluster_kernel
    if (g == 0 && p.strict_out) p.strict_out[scene] = 1;       // the proof implies strict unique maxima
    int32_t *oidx = p.idx + (size_t)scene * p.npoint;
    for (unsigned j = g; j < (unsigned)p.npoint; j += C * THREADS) oidx[j] = (int)j;
    if (p.new_xyz) {
      float *o = p.new_xyz + (size_t)scene * p.npoint * 3;
      for (unsigned e = g; e < 3u * (unsigned)p.npoint; e += C * THREADS) o[e] = __ldg(xyz + e);
    }
    return;
  }

  // ---- load: (v, k) of my slots into smem, insertion-sort them by v, then fetch the coordinates ---
  int nvalid = 0;
  for (int s = 0; s < P; ++s) {
    const unsigned q = q0 + s;
    unsigned v = 0xffffffffu;
    int k = 0;
    if (q < (unsigned)p.N) {
      k = __ldg(perm + q);
      v = v_of_k((unsigned)k);
      ++nvalid;
    }
    int pos = s;                                                 // insertion sort (ascending v)
    while (pos > 0 && sv[(pos - 1) * THREADS + tid] > v) {
      sv[pos * THREADS + tid] = sv[(pos - 1) * THREADS + tid];
      sk[pos * THREADS + tid] = sk[(pos - 1) * THREADS + tid];
      --pos;
    }
    sv[pos * THREADS + tid] = v;
    sk[pos * THREADS + tid] = k;
  }
  float x[P], y[P], z[P], t[P];
  int kk[P];
#pragma unroll
  for (int s = 0; s < P; ++s) kk[s] = sk[s * THREADS + tid];   // before the planes are overwritten
  float lox = INFINITY, loy = INFINITY, loz = INFINITY, hix = -INFINITY, hiy = -INFINITY, hiz = -INFINITY;
#pragma unroll
  for (int s = 0; s < P; ++s) {
    float px = 0.f, py = 0.f, pz = 0.f, pt = -1.0f;
    if (s < nvalid) {
      const int k = kk[s];
      px = __ldg(xyz + 3 * k + 0);
      py = __ldg(xyz + 3 * k + 1);
      pz = __ldg(xyz + 3 * k + 2);
      const float mag = __fmaf_rn(pz, pz, __fmaf_rn(px, px, __fmul_rn(py, py)));
      pt = ((double)mag <= 1e-3) ? -1.0f : 1e10f;   // reference compares in double (F5)
      if (pt > 0.f) {
        lox = fminf(lox, px); loy = fminf(loy, py); loz = fminf(loz, pz);
        hix = fmaxf(hix, px); hiy = fmaxf(hiy, py); hiz = fmaxf(hiz, pz);
      }
    }
    sx[s * THREADS + tid] = px;
    sy[s * THREADS + tid] = py;
    sz[s * THREADS + tid] = pz;
    if (XYZ_SMEM) sk16[s * THREADS + tid] = (uint16_t)kk[s];
    x[s] = px; y[s] = py; z[s] = pz;
    t[s] = pt;
  }
  // bounding box of the warp's selectable points
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o)); hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
    loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o)); hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
    loz = fminf(loz, __shfl_xor_sync(0xffffffffu, loz, o)); hiz = fmaxf(hiz, __shfl_xor_sync(0xffffffffu, hiz, o));
  }
  const float p0x = __ldg(xyz + 0), p0y = __ldg(xyz + 1), p0z = __ldg(xyz + 2);
  float ox = p0x, oy = p0y, oz = p0z;
  int32_t *idx = p.idx + (size_t)scene * p.npoint;
  float *nxyz = p.new_xyz ? p.new_xyz + (size_t)scene * p.npoint * 3 : nullptr;
  const bool writer = (rank == 0 && tid == THREADS - 1);
  if (writer) {
    idx[0] = 0;
    if (nxyz) { nxyz[0] = ox; nxyz[1] = oy; nxyz[2] = oz; }
    if (TRACK && p.strict_out) p.strict_out[scene] = 1;   // cleared by whoever sees a tie (ordered by the sync below)
  }
  const unsigned tx_bytes = 24u * C;
  uint32_t r_slot0 = 0, r_slot1 = 0, r_bar0 = 0, r_bar1 = 0;
  if (C > 1) {
    if (tid == 0) {
      fps_mbar_init(&c_bar[0], 1);
      fps_mbar_init(&c_bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      fps_mbar_arm(&c_bar[0], tx_bytes);
      fps_mbar_arm(&c_bar[1], tx_bytes);
    }
    cluster.sync();
    if (warp == 0 && lane < C) {
      r_slot0 = mapa_cluster(fps_s2u(&c_cand[0][rank]), lane);
      r_slot1 = mapa_cluster(fps_s2u(&c_cand[1][rank]), lane);
      r_bar0 = mapa_cluster(fps_s2u(&c_bar[0]), lane);
      r_bar1 = mapa_cluster(fps_s2u(&c_bar[1]), lane);
    }
  } else {
    __syncthreads();
  }
  // The warp's candidate lives in w_cand (both buffers once it is stable); only its key is kept in a
  // register for the cull test.  key 0 = nothing selectable.
  unsigned ck = 0u;
  bool fresh = false;                            // the first round must compute
  bool stale = false;                            // the OTHER buffer still holds an older candidate
  int my_k = -1;                                 // point index of the warp's candidate
  int old = -2;                                  // the last pick

  // arg-max over the lanes holding (key, v): largest key, then smallest v.  The common case (a unique
  // maximum) costs two dependent redux; the ballot that detects ties issues alongside the second.
  auto argmax_lane = [&](bool valid, unsigned key, unsigned v, unsigned &kmax, bool &tied) -> unsigned {
    kmax = __reduce_max_sync(0xffffffffu, valid ? key : 0u);
    const bool hit = valid && key == kmax;
    unsigned src = __reduce_min_sync(0xffffffffu, hit ? lane : 32u);
    const unsigned ties = __ballot_sync(0xffffffffu, hit);
    tied = kmax != 0u && (ties & (ties - 1u)) != 0u;
    if (tied) src = fps_resolve_tie(hit, v, lane);   // rare
    return src & 31u;
  };
  // "Every pick so far was the strict unique maximum" (p.strict_out): then FPS over any prefix of the OUTPUT is
  // the identity, which lets the next set-abstraction layers skip their sampling without the proof kernels.
  // Kept off the round's dependency chain: ties between warps / CTAs are seen by every thread (`strict`), ties
  // inside the winning warp or thread are checked lazily by the ONE thread that owns the round's winner, which
  // then stores 0 (the flag was initialised to 1 before the first round).
  bool strict = true;        // uniform: no tie at CTA / cluster level so far, and every round had a candidate
  bool own_r = false;        // this lane published the warp's current candidate ...
  bool wt_r = false;         // ... which tied with another lane of the warp
  float best_r = -1.0f;      // this thread's best min-distance at its last update
  int mk_r = -1;             // and the point that holds it

  for (int j = 1; j < p.npoint; ++j) {
    const int buf = j & 1;
    // ---- can this centre change any min-distance of the warp? -------------------------------------
    // (the warp whose own candidate was just picked certainly must update: no test on the critical path)
    bool skip = false;
    if (fresh && old != my_k) {
      const float ex = fmaxf(0.f, fmaxf(lox - ox, ox - hix)), ey = fmaxf(0.f, fmaxf(loy - oy, oy - hiy)),
                  ez = fmaxf(0.f, fmaxf(loz - oz, oz - hiz));
      const float lb2 = (ex * ex + ey * ey + ez * ez) * 0.99999f;
      skip = ck == 0u ? !(hix >= lox) : lb2 >= __uint_as_float(ck - 1u);
    }
    if (!skip) {
      float bvv[P];
      int bsi[P];
#pragma unroll
      for (int s = 0; s < P; ++s) {
        const float d2 = XYZ_SMEM ? fminf(sqdist_ref(sx[s * THREADS + tid], sy[s * THREADS + tid],
                                                     sz[s * THREADS + tid], ox, oy, oz), t[s])
                                  : fminf(sqdist_ref(x[s], y[s], z[s], ox, oy, oz), t[s]);
        t[s] = d2;
        bvv[s] = d2;
        bsi[s] = s;
      }
      // tournament over the slots (depth log2 P instead of a P-long dependent chain); the left operand
      // always covers the lower slots, so strict '>' keeps the lowest slot (= lowest v) among equals
#pragma unroll
      for (int stride = 1; stride < P; stride *= 2) {
#pragma unroll
        for (int s = 0; s + stride < P; s += 2 * stride) {
          if (bvv[s + stride] > bvv[s]) { bvv[s] = bvv[s + stride]; bsi[s] = bsi[s + stride]; }
        }
      }
      const float best = bvv[0];
      const int bs = bsi[0];
      const float mx = sx[bs * THREADS + tid], my = sy[bs * THREADS + tid], mz = sz[bs * THREADS + tid];
      const int mk = XYZ_SMEM ? (int)sk16[bs * THREADS + tid] : sk[bs * THREADS + tid];
      const unsigned mv = XYZ_SMEM ? v_of_k((unsigned)mk) : sv[bs * THREADS + tid];
      const unsigned key = best < 0.f ? 0u : __float_as_uint(best) + 1u;
      bool wtied;
      const unsigned wsrc = argmax_lane(true, key, mv, ck, wtied);
      if (lane == wsrc) {
        FpsCandV *e = &w_cand[buf][warp];
        *reinterpret_cast<uint4 *>(e) = make_uint4(ck, mv, (unsigned)mk, __float_as_uint(mx));
        *reinterpret_cast<float2 *>(&e->y) = make_float2(my, mz);
      }
      my_k = ck ? __shfl_sync(0xffffffffu, mk, wsrc) : -1;   // consumed next round only
      if (TRACK) { own_r = lane == wsrc; wt_r = wtied; best_r = best; mk_r = mk; }
      fresh = true;
      stale = true;
    } else if (stale) {
      if (lane == 0) {
        const FpsCandV *src = &w_cand[buf ^ 1][warp];
        FpsCandV *dst = &w_cand[buf][warp];
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(src);
        *reinterpret_cast<float2 *>(&dst->y) = *reinterpret_cast<const float2 *>(&src->y);
      }
      stale = false;
    }
    __syncthreads();
    // ---- CTA arg-max, computed redundantly by every warp ------------------------------------------
    uint4 e4 = make_uint4(0u, 0xffffffffu, 0u, 0u);
    float2 eyz = make_float2(0.f, 0.f);
    if (lane < NW) {
      e4 = *reinterpret_cast<const uint4 *>(&w_cand[buf][lane]);
      eyz = *reinterpret_cast<const float2 *>(&w_cand[buf][lane].y);
    }
    unsigned bmax;
    bool ctied;
    unsigned src = argmax_lane(lane < NW, e4.x, e4.y, bmax, ctied);
    unsigned bv = __shfl_sync(0xffffffffu, e4.y, src);
    unsigned wk = __shfl_sync(0xffffffffu, e4.z, src);
    if (TRACK) strict = strict && !ctied;
    unsigned wxb = __shfl_sync(0xffffffffu, e4.w, src);
    float wy = __shfl_sync(0xffffffffu, eyz.x, src);
    float wz = __shfl_sync(0xffffffffu, eyz.y, src);
    if (C > 1) {
      if (warp == 0 && lane < C) {
        const uint32_t rs = buf ? r_slot1 : r_slot0, rb = buf ? r_bar1 : r_bar0;
        st_async_v4(rs, bmax, bv, wk, wxb, rb);
        st_async_v2(rs + 16, __float_as_uint(wy), __float_as_uint(wz), rb);
      }
      fps_mbar_wait(&c_bar[buf], (unsigned)((j - 1) >> 1) & 1u);
      if (tid == 0) fps_mbar_arm(&c_bar[buf], tx_bytes);
      e4 = make_uint4(0u, 0xffffffffu, 0u, 0u);
      eyz = make_float2(0.f, 0.f);
      if (lane < C) {
        e4 = *reinterpret_cast<const uint4 *>(&c_cand[buf][lane]);
        eyz = *reinterpret_cast<const float2 *>(&c_cand[buf][lane].y);
      }
      src = argmax_lane(lane < C, e4.x, e4.y, bmax, ctied);
      wk = __shfl_sync(0xffffffffu, e4.z, src);
      if (TRACK) strict = strict && !ctied;
      wxb = __shfl_sync(0xffffffffu, e4.w, src);
      wy = __shfl_sync(0xffffffffu, eyz.x, src);
      wz = __shfl_sync(0xffffffffu, eyz.y, src);
    }
    old = 0;
    if (TRACK) strict = strict && bmax != 0u;
    if (bmax == 0u) { ox = p0x; oy = p0y; oz = p0z; }
    else { ox = __uint_as_float(wxb); oy = wy; oz = wz; old = (int)wk; }
    if (TRACK && own_r && bmax != 0u && old == mk_r) {   // one thread of the cluster per round
      int eq = 0;
#pragma unroll
      for (int s = 0; s < P; ++s) eq += (t[s] == best_r) ? 1 : 0;
      if (wt_r || eq > 1) p.strict_out[scene] = 0;
    }
    if (writer) {
      idx[j] = old;
      if (nxyz) { nxyz[3 * j + 0] = ox; nxyz[3 * j + 1] = oy; nxyz[3 * j + 2] = oz; }
    }
  }
  if (writer && p.strict_out && (!TRACK || !strict)) p.strict_out[scene] = 0;
  if (C > 1) cluster.sync();
}

template <int P, int THREADS, bool XYZ_SMEM, bool TRACK>
static int launch_fps_cull_t(const FpsParams &p, const int32_t *perm, int B, int C, cudaStream_t stream) {
  auto kern = fps_cull_kernel<P, THREADS, XYZ_SMEM, TRACK>;
  const size_t smem = XYZ_SMEM ? (size_t)P * THREADS * (3 * sizeof(float) + sizeof(uint16_t))
                               : (size_t)5 * P * THREADS * sizeof(float);
  static bool attr_set = false;   // per instantiation; the attribute is sticky for the process
  if (!attr_set) {
    SPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C, B, 1);
  cfg.blockDim = dim3(THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SPC_CUDA(cudaLaunchKernelEx(&cfg, kern, p, perm));
  return SPC_OK;
}

template <int P, int THREADS, bool XYZ_SMEM>
static int launch_fps_cull(const FpsParams &p, const int32_t *perm, int B, int C, cudaStream_t stream) {
  return p.strict_out ? launch_fps_cull_t<P, THREADS, XYZ_SMEM, true>(p, perm, B, C, stream)
                      : launch_fps_cull_t<P, THREADS, XYZ_SMEM, false>(p, perm, B, C, stream);
}

// ------------------------------------------------------------------------------------------------
// "Verify in parallel what would be constructed sequentially".
// SA2..SA4 of the detector run FPS on the OUTPUT of the previous FPS (an FPS-ordered point list),
// where the answer is 0..npoint-1 unless exact distance ties interfere (the reference
// model silently relies on this).  Whether FPS(xyz)[0:npoint] == identity can be CHECKED with no
// sequential dependency: with the first j points selected, point j must be the unique maximum of
// the running min-distance among all not-yet-selected points.  Kernel A computes
// D[j] = min_{i<j} d(p_j, p_i) (what round j's winner scored); kernel B recomputes every point's
// running min-distance and flags any k > j that reaches D[j] (a tie or a larger value -- then
// the tie-break / order decides and the full kernel must run).  Same fp32 arithmetic as the
// sequential kernel (sqdist_ref, fminf, temp = 1e10), so the proof is exact, not approximate.
// Cost: N*npoint distance evaluations, fully parallel (~10 us for 2048 -> 1024 at B=8) instead
// of npoint-1 sequential rounds (~290 us).
// ------------------------------------------------------------------------------------------------
__global__ void fps_fill_kernel(int *p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

constexpr int CHK_THREADS = 128;
constexpr int CHK_TILE = 256;

__device__ __forceinline__ bool fps_skipped(float x, float y, float z) {
  const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
  return (double)mag <= 1e-3;
}

// D[b][j] = min(1e10, min_{i<j} d(p_j, p_i)) for j < npoint;   ok[b] = 0 if a prefix point other than
// p_0 is one the reference never selects (|p|^2 <= 1e-3)
__global__ void __launch_bounds__(CHK_THREADS) fps_prefix_dist_kernel(const float *__restrict__ xyz, int N,
                                                                       int npoint, float *__restrict__ D,
                                                                       int *__restrict__ ok,
                                                                       const int *__restrict__ known) {
  __shared__ float sx[CHK_TILE], sy[CHK_TILE], sz[CHK_TILE];
  const int b = blockIdx.y;
  if (known != nullptr && known[b] != 0) return;   // already established by the call that produced xyz
  const float *P = xyz + (size_t)b * N * 3;
  const int j = blockIdx.x * CHK_THREADS + threadIdx.x;
  const bool act = j < npoint;
  const float px = act ? __ldg(P + 3 * j) : 0.f, py = act ? __ldg(P + 3 * j + 1) : 0.f,
              pz = act ? __ldg(P + 3 * j + 2) : 0.f;
  if (act && j > 0 && fps_skipped(px, py, pz)) ok[b] = 0;
  float m = 1e10f;
  const int jmax = min(npoint, (int)(blockIdx.x + 1) * CHK_THREADS);   // largest j of this block + 1
  for (int base = 0; base < jmax - 1; base += CHK_TILE) {
    const int tile = min(CHK_TILE, jmax - 1 - base);
    __syncthreads();
    for (int e = threadIdx.x; e < tile * 3; e += CHK_THREADS) {
      const float v = __ldg(P + (size_t)base * 3 + e);
      const int pt = e / 3, c = e - pt * 3;
      (c == 0 ? sx : c == 1 ? sy : sz)[pt] = v;
    }
    __syncthreads();
    const int lim = min(tile, j - base);          // only i < j
    for (int i = 0; i < lim; ++i) {
      // the reference never updates temp from a skipped centre?  No: the CENTRE may be any point
      // (only index 0 can be a skipped one, and it is used as a centre like any other)
      m = fminf(sqdist_ref(px, py, pz, sx[i], sy[i], sz[i]), m);
    }
  }
  if (act) {
    D[(size_t)b * npoint + j] = m;
    // a winner at distance 0 (duplicate of an earlier point) ties with every selected point
    if (j > 0 && !(m > 0.f)) ok[b] = 0;
  }
}

// ok[b] &= for all rounds j in [1, npoint) and all points k > j (not skipped):
//            min(1e10, min_{i<j} d(p_k, p_i)) < D[j]
__global__ void __launch_bounds__(CHK_THREADS) fps_prefix_check_kernel(const float *__restrict__ xyz, int N,
                                                                        int npoint,
                                                                        const float *__restrict__ D,
                                                                        int *__restrict__ ok,
                                                                        const int *__restrict__ known) {
  __shared__ float sx[CHK_TILE], sy[CHK_TILE], sz[CHK_TILE], sd[CHK_TILE];
  const int b = blockIdx.y;
  if (known != nullptr && known[b] != 0) return;
  const float *P = xyz + (size_t)b * N * 3;
  const float *Db = D + (size_t)b * npoint;
  const int k = blockIdx.x * CHK_THREADS + threadIdx.x;
  const bool act = k < N;
  const float px = act ? __ldg(P + 3 * k) : 0.f, py = act ? __ldg(P + 3 * k + 1) : 0.f,
              pz = act ? __ldg(P + 3 * k + 2) : 0.f;
  const bool cand = act && !fps_skipped(px, py, pz);   // skipped points are never candidates
  float m = 1e10f;
  bool bad = false;
  for (int base = 0; base < npoint - 1; base += CHK_TILE) {
    const int tile = min(CHK_TILE, npoint - 1 - base);
    __syncthreads();
    for (int e = threadIdx.x; e < tile * 3; e += CHK_THREADS) {
      const float v = __ldg(P + (size_t)base * 3 + e);
      const int pt = e / 3, c = e - pt * 3;
      (c == 0 ? sx : c == 1 ? sy : sz)[pt] = v;
    }
    for (int e = threadIdx.x; e < tile; e += CHK_THREADS) sd[e] = __ldg(Db + base + e + 1);   // D[j], j = i+1
    __syncthreads();
    for (int i = 0; i < tile; ++i) {
      m = fminf(sqdist_ref(px, py, pz, sx[i], sy[i], sz[i]), m);
      // round j = base+i+1 picks among points with index > j - 1 that are not yet selected
      bad |= (k > base + i + 1) && (m >= sd[i]);
    }
  }
  if (cand && bad) ok[b] = 0;   // benign race: every writer stores 0
}

template <int P, int THREADS, bool XYZ_REGS>
static int launch_fps(const FpsParams &p, int B, int C, cudaStream_t stream) {
  auto kern = fps_cluster_kernel<P, THREADS, XYZ_REGS>;
  const size_t smem = (size_t)4 * P * THREADS * sizeof(float);
  SPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (C > 8) SPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C, B, 1);
  cfg.blockDim = dim3(THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SPC_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return SPC_OK;
}

// how many clusters of size C (512 threads, given smem) can be co-resident; cached per C
template <int P, int THREADS, bool XYZ_REGS>
static int max_active_clusters(int C) {
  auto kern = fps_cluster_kernel<P, THREADS, XYZ_REGS>;
  const size_t smem = (size_t)4 * P * THREADS * sizeof(float);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (C > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C, 1, 1);
  cfg.blockDim = dim3(THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  return n;
}

}  // namespace spc

using namespace spc;

#define FPS_CASE(PV, TH, REGS)                                         \
  case PV: return launch_fps<PV, TH, REGS>(p, B, C, stream);

static int fps_impl(const float *xyz, int B, int N, int npoint, int32_t *idx, float *new_xyz,
                    int hint_ordered, const int32_t *known_ordered, int32_t *strict_out, void *workspace,
                    size_t workspace_bytes, void *stream_);

// process-wide tuning knob (0 = automatic).  8-CTA clusters minimise the latency of one call;
// 4-CTA clusters cost ~15 % more time per call but half the SM-time, which is what matters when
// several batches are in flight on different streams (spacap3d_b200/pipeline.py).
static int g_fps_cluster = 0;
extern "C" int spc_set_fps_cluster(int cluster_ctas) {
  if (cluster_ctas != 0 && cluster_ctas != 1 && cluster_ctas != 2 && cluster_ctas != 4 && cluster_ctas != 8 &&
      cluster_ctas != 16) {
    set_error("spc_set_fps_cluster: %d is not one of 0,1,2,4,8,16", cluster_ctas);
    return SPC_ERR_INVALID_ARG;
  }
  g_fps_cluster = cluster_ctas;
  return SPC_OK;
}

// process-wide switch for the Morton-sorted, culled kernel (0 = off, 1 = on).  Measured on B200, batch
// 8 x 40 000 -> 2048: one call alone is ~15 % slower with it (0.67 vs 0.60 us per round: the explicit
// tie-break and the cull test sit on the round's dependency chain, and the sort prepass is extra), but it
// issues a fraction of the instructions, so a pipeline that keeps several batches in flight on different
// streams gains ~5 % overall.  spacap3d_b200/pipeline.py turns it on; SPC_FPS_CULL=0/1 overrides.
static int g_fps_cull = 0;
extern "C" int spc_set_fps_cull(int on) {
  if (on != 0 && on != 1 && on != 2 && on != 3) {
    set_error("spc_set_fps_cull: %d is not 0, 1, 2 or 3", on);
    return SPC_ERR_INVALID_ARG;
  }
  g_fps_cull = on;
  return SPC_OK;
}

extern "C" size_t spc_fps_workspace_bytes(int B, int N, int npoint) {
  // D (B,npoint) + ok (B) for the ordered-prefix proof, perm (B,N) for the Morton-sorted (culled) kernel
  return ((size_t)B * (size_t)(npoint > 0 ? npoint : 0) + (size_t)B + (size_t)B * (size_t)(N > 0 ? N : 0)) * 4;
}

extern "C" int spc_furthest_point_sampling(const float *xyz, int B, int N, int npoint,
                                           int32_t *idx, float *new_xyz, void *stream_) {
  return fps_impl(xyz, B, N, npoint, idx, new_xyz, 0, nullptr, nullptr, nullptr, 0, stream_);
}

extern "C" int spc_furthest_point_sampling_ex(const float *xyz, int B, int N, int npoint,
                                              int32_t *idx, float *new_xyz, int hint_ordered,
                                              void *workspace, size_t workspace_bytes,
                                              void *stream_) {
  return fps_impl(xyz, B, N, npoint, idx, new_xyz, hint_ordered, nullptr, nullptr, workspace, workspace_bytes, stream_);
}

extern "C" int spc_furthest_point_sampling_ex2(const float *xyz, int B, int N, int npoint, int32_t *idx,
                                               float *new_xyz, int hint_ordered, const int32_t *known_ordered,
                                               int32_t *strict_out, void *workspace, size_t workspace_bytes,
                                               void *stream_) {
  return fps_impl(xyz, B, N, npoint, idx, new_xyz, hint_ordered, known_ordered, strict_out, workspace,
                  workspace_bytes, stream_);
}

static int fps_cull_mode() {
  if (const char *e = getenv("SPC_FPS_CULL")) return atoi(e);
  return g_fps_cull;
}
static bool fps_cull_enabled() { return fps_cull_mode() != 0; }

static int fps_impl(const float *xyz, int B, int N, int npoint, int32_t *idx, float *new_xyz,
                    int hint_ordered, const int32_t *known_ordered, int32_t *strict_out, void *workspace,
                    size_t workspace_bytes, void *stream_) {
  SPC_CHECK_ARG(B >= 0 && N >= 1 && npoint >= 0, "fps: bad sizes B=%d N=%d npoint=%d", B, N, npoint);
  SPC_CHECK_ARG(xyz && (idx || npoint == 0 || B == 0), "fps: null pointer");
  if (B == 0 || npoint == 0) return SPC_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  FpsParams p;
  p.xyz = xyz; p.idx = idx; p.new_xyz = new_xyz; p.N = N; p.npoint = npoint;
  p.T = ref_opt_n_threads(N);
  p.log2T = 0;
  while ((1 << p.log2T) < p.T) ++p.log2T;
  p.Q = (N + p.T - 1) / p.T;
  p.ordered_ok = nullptr;
  p.strict_out = strict_out;
  // ---- optional verified shortcut for FPS-ordered inputs ---------------------------------------
  if (hint_ordered && workspace && npoint >= 2 && npoint <= N &&
      workspace_bytes >= spc_fps_workspace_bytes(B, N, npoint) && B <= 65535 &&
      (long long)N * npoint <= (1LL << 26)) {
    float *D = reinterpret_cast<float *>(workspace);
    int *ok = reinterpret_cast<int *>(D + (size_t)B * npoint);
    fps_fill_kernel<<<ceil_div(B, 128), 128, 0, stream>>>(ok, B, 1);
    fps_prefix_dist_kernel<<<dim3(ceil_div(npoint, CHK_THREADS), B), CHK_THREADS, 0, stream>>>(xyz, N, npoint, D, ok, known_ordered);
    fps_prefix_check_kernel<<<dim3(ceil_div(N, CHK_THREADS), B), CHK_THREADS, 0, stream>>>(xyz, N, npoint, D, ok, known_ordered);
    SPC_LAUNCH_CHECK("fps_prefix_check");
    p.ordered_ok = ok;
  }
  if (p.ordered_ok == nullptr && known_ordered != nullptr && npoint <= N) p.ordered_ok = known_ordered;
  const long long V = (long long)p.T * p.Q;

  // ---- small clouds: one CTA of 256 threads -------------------------------------------------
  if (V <= 256 * 16) {
    const int C = 1;
    const int need = (int)((V + 255) / 256);
    const int P = need <= 1 ? 1 : need <= 2 ? 2 : need <= 4 ? 4 : need <= 8 ? 8 : 16;
    switch (P) {
      FPS_CASE(1, 256, true) FPS_CASE(2, 256, true) FPS_CASE(4, 256, true)
      FPS_CASE(8, 256, true) FPS_CASE(16, 256, true)
    }
  }
  // ---- clusters of 512-thread CTAs ------------------------------------------------------------
  // The rounds are latency-bound, so all B scenes should be co-resident.  Measured on B200
  // (B=8, N=40k): C=8 0.81 us/round, C=16 1.03 (slower exchange across a non-portable cluster),
  // C=4 1.19 (xyz no longer fits in registers) -> prefer 8; SPC_FPS_CLUSTER overrides for tuning.
  static int cached_max16 = -1, cached_max8 = -1;
  if (cached_max8 < 0) cached_max8 = max_active_clusters<10, 512, true>(8);
  int C = 8;
  if (cached_max8 > 0 && cached_max8 < B) C = (B * 4 <= kNumSMs) ? 4 : 2;
  if (g_fps_cluster) C = g_fps_cluster;
  if (const char *e = getenv("SPC_FPS_CLUSTER")) { int c = atoi(e); if (c == 1 || c == 2 || c == 4 || c == 8 || c == 16) C = c; }
  int need = (int)((V + (long long)C * 512 - 1) / ((long long)C * 512));
  while (need > 27 && C < 16) { C *= 2; need = (int)((V + (long long)C * 512 - 1) / ((long long)C * 512)); }
  if (C == 16 && cached_max16 < 0) cached_max16 = max_active_clusters<6, 512, true>(16);
  if (C == 16 && cached_max16 <= 0) {
    set_error("fps: N=%d needs a 16-CTA cluster which this device cannot schedule", N);
    return SPC_ERR_UNSUPPORTED;
  }
  if (need > 27) {
    set_error("fps: N=%d exceeds the on-chip capacity of a 16-CTA cluster (max %d points)", N, 16 * 512 * 27);
    return SPC_ERR_UNSUPPORTED;
  }
  while (C > 1 && need <= 1) { C /= 2; need = (int)((V + (long long)C * 512 - 1) / ((long long)C * 512)); }
  // ---- Morton-sorted, culled kernel: needs the caller's workspace (perm) ----------------------------
  if (workspace && workspace_bytes >= spc_fps_workspace_bytes(B, N, npoint) && N >= 8192 && npoint >= 64 &&
      B <= 65535 && C <= 8 && fps_cull_enabled()) {
    // mode 3: "full-SM" CTAs -- 768 threads x 20 points, ~215 KB of shared memory, clusters of ceil(N / 15360).
    // The same work as mode 2 (three 256-thread CTAs per SM) but packed by construction: a 40 k-point scene holds
    // exactly 3 SMs.  The hardware spreads the 64 small CTAs of a mode-2 batch over up to 64 SMs, where each of them
    // blocks kernels that need a whole SM (the fused SA kernel) for the 1.5 ms the sampler runs.
    const int cull_mode = fps_cull_mode();
    const int TH = cull_mode == 3 ? 768 : 256;
    const int Cc = cull_mode == 3 ? (int)(((long long)N + 768LL * 20 - 1) / (768LL * 20)) : C;
    const int need256 = (int)(((long long)N + (long long)Cc * TH - 1) / ((long long)Cc * TH));
    if (need256 <= 20 && Cc <= 8 && N <= FMS_THREADS * FMS_ITEMS) {
      int32_t *perm = reinterpret_cast<int32_t *>(workspace) + (size_t)B * npoint + (size_t)B;
      const size_t sort_smem = (size_t)FMS_BINS * sizeof(int);
      static bool sort_attr_set = false;
      if (!sort_attr_set) {
        SPC_CUDA(cudaFuncSetAttribute(fps_morton_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));
        sort_attr_set = true;
      }
      fps_morton_sort_kernel<<<B, FMS_THREADS, sort_smem, stream>>>(xyz, N, perm);
      SPC_LAUNCH_CHECK("fps_morton_sort_kernel");
      const int P256 = need256 <= 8 ? 8 : need256 <= 10 ? 10 : need256 <= 16 ? 16 : 20;
      if (cull_mode == 3) {
        switch (P256) {
          case 8: return launch_fps_cull<8, 768, true>(p, perm, B, Cc, stream);
          case 10: return launch_fps_cull<10, 768, true>(p, perm, B, Cc, stream);
          case 16: return launch_fps_cull<16, 768, true>(p, perm, B, Cc, stream);
          default: return launch_fps_cull<20, 768, true>(p, perm, B, Cc, stream);
        }
      }
      const bool xyz_smem = cull_mode == 2;
      switch (P256) {
        case 8: return xyz_smem ? launch_fps_cull<8, 256, true>(p, perm, B, C, stream) : launch_fps_cull<8, 256, false>(p, perm, B, C, stream);
        case 10: return xyz_smem ? launch_fps_cull<10, 256, true>(p, perm, B, C, stream) : launch_fps_cull<10, 256, false>(p, perm, B, C, stream);
        case 16: return xyz_smem ? launch_fps_cull<16, 256, true>(p, perm, B, C, stream) : launch_fps_cull<16, 256, false>(p, perm, B, C, stream);
        default: return xyz_smem ? launch_fps_cull<20, 256, true>(p, perm, B, C, stream) : launch_fps_cull<20, 256, false>(p, perm, B, C, stream);
      }
    }
  }
  // 256-thread CTAs with 20 points per thread and TWO CTAs per SM whenever the cloud fits: fewer
  // warps per reduction level make a round faster (1.32 vs 1.51 ms for 40k -> 2048 at batch 8) and
  // two latency-bound CTAs (of different scenes / batches) share one SM's issue slots, which halves
  // the SM-time per scene (measured: 9.4k vs 7.2k scenes/s with 8 batches in flight).
  // SPC_FPS_THREADS=512 restores the one-CTA-per-SM variant for A/B runs.
  {
    const char *e = getenv("SPC_FPS_THREADS");
    const int need256 = (int)((V + (long long)C * 256 - 1) / ((long long)C * 256));
    if (!(e && atoi(e) == 512) && need256 <= 20 && need256 > 4) {
      const int P256 = need256 <= 8 ? 8 : need256 <= 10 ? 10 : need256 <= 16 ? 16 : 20;
      switch (P256) {
        FPS_CASE(8, 256, true) FPS_CASE(10, 256, true) FPS_CASE(16, 256, true) FPS_CASE(20, 256, true)
      }
    }
  }
  // P = 27 is the most that fits: 4 arrays x 27 x 512 x 4 B = 216 KB of the 227 KB a CTA may use
  static const int opts[] = {2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 27};
  int P = 27;
  for (int o : opts) if (o >= need) { P = o; break; }
  switch (P) {
    FPS_CASE(2, 512, true) FPS_CASE(3, 512, true) FPS_CASE(4, 512, true) FPS_CASE(5, 512, true)
    FPS_CASE(6, 512, true) FPS_CASE(8, 512, true) FPS_CASE(10, 512, true) FPS_CASE(12, 512, true)
    FPS_CASE(16, 512, true) FPS_CASE(20, 512, true) FPS_CASE(24, 512, false) FPS_CASE(27, 512, false)
  }
  set_error("fps: internal dispatch error");
  return SPC_ERR_UNSUPPORTED;
}
