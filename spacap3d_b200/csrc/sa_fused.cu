// sa_fused.cu -- fused set-abstraction forward (eval mode): grouping + relative-xyz
// normalisation + shared MLP (3 x [1x1 conv + folded BN + ReLU]) + max-pool over nsample in ONE
// kernel, the two wide 1x1 convs on tcgen05 tensor cores with TMEM accumulators (sm_100a).
//
// Replaces, for PointnetSAModuleVotes.forward in eval mode (reference pointnet2_modules.py:244-271):
//   QueryAndGroup's two group_points launches + sub + div + cat (pointnet2_utils.py:351-362),
//   SharedMLP = 3 x (cuDNN conv, cuDNN BN, ReLU) (pytorch_utils.py:11-36) and F.max_pool2d --
// i.e. ~14 kernels that each stream a (B, C, npoint, nsample) activation through HBM
// (~270 MB/scene unfused vs ~8 MB compulsory).
//
// Algorithm (per tile of 128 rows, a row = one (centre, neighbour) pair):
//   layer 0  h1 = relu(W0' . [ (p_i - c_j)/r , f_i ] + b0)         CUDA cores, in the gather stage
//            MODE_PROJ: the feature part of conv0 only depends on the POINT, not on the pair, so
//            it is hoisted out of the grouping: G[i] = W0' . [p_i/r, f_i] is precomputed per point
//            (npoint*nsample/n = 4..16x fewer MACs) and Hc[j] = b0 - W0x' . c_j/r per centre; the
//            stage then only gathers:  h1 = relu(G[idx] + Hc[j]).
//            MODE_INLINE (few input channels, SA1): evaluated directly from xyz and raw features.
//   layer 1  D1[128 rows x C2]  = H1[128 x C1] . W1'^T             tcgen05.mma, M=128, N=C2
//            h2 = relu(D1 + b1) -> bf16 -> shared memory (thread per row, TMEM lane = row)
//   layer 2  D2[C3 x 128 rows]  = W2'[C3 x C2] . H2^T              tcgen05.mma, TRANSPOSED so that a
//            TMEM lane is an output CHANNEL and the 128 columns are the rows of the tile: the
//            max over the nsample neighbours of a centre is then a register-only reduction.
//   out[b, c, j] = relu(max_k D2[c, j*ns+k] + b2[c])   (ReLU and +b commute with max)
// BN (eval) is folded on the host: W' = diag(gamma/sqrt(var+eps)) W, b = beta - mean*scale.
//
// Shared-memory operand layout: canonical UMMA K-major, no swizzle: 8-element (16 B) chunks,
// element (row, k) at (k/8)*ROWS*16 + row*16 + (k%8)*2 bytes, i.e. core matrices of 8 rows x 16 B
// are contiguous (SBO = 128 B) and K-chunks are ROWS*16 B apart (LBO).  Both the gather stage and
// the epilogue write one 16-byte chunk per lane with lane = row, which is bank-conflict free.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace spc {

// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, tcgen05 (alloc / mma / commit / ld / fences)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s2u(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarrier_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarrier_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SA_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SA_DONE;\n"
      "bra SA_WAIT;\n"
      "SA_DONE:\n"
      "}\n" ::"r"(s2u(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T ; one thread issues for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(bar))
               : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading (K-direction) byte offset >> 4 |
//   [32,46) stride (8-row group) byte offset >> 4 | [46,48) version = 1 | [61,64) layout = 0
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes, int swap = 0) {
  if (swap) { const uint32_t t = lbo_bytes; lbo_bytes = sbo_bytes; sbo_bytes = t; }
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): c=F32, a=b=BF16, both K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}

// ---------------------------------------------------------------------------------------------
struct SaFusedParams {
  const float *xyz;       // (B,n,3)
  const float *new_xyz;   // (B,np,3)
  const int32_t *idx;     // (B,np,ns)
  // MODE_PROJ
  const float *G;         // (B,n,C1)  per-point projection (BN scale folded)
  const float *Hc;        // (B,np,C1) per-centre bias
  // MODE_INLINE
  const float *feat;      // (B,Cf,n) or nullptr
  const float *W0;        // (C1, 3+Cf) folded, fp32
  const float *b0;        // (C1)
  int Cf;
  float radius;           // divide relative xyz by this (1.0 when normalize_xyz is off)
  const __nv_bfloat16 *W1;  // (C2,C1) folded, bf16 row-major
  const float *b1;          // (C2)
  const __nv_bfloat16 *W2;  // (C3,C2)
  const float *b2;          // (C3)
  float *out;             // (B,C3,np)
  int B, n, np, ns;
  int num_tiles;          // B*np*ns/128
  int desc_swap;          // debug (env SPC_SA_DESC_SWAP=1): swap LBO/SBO roles in the descriptors
};

constexpr int SA_THREADS = 256;
constexpr int SA_ROWS = 128;        // rows (centre,neighbour pairs) per tile
constexpr int SA_MAX_K0 = 3 + 16;   // MODE_INLINE supports up to 16 raw feature channels
constexpr int SA_W0_STRIDE = 20;    // floats per channel row of the inline layer-0 weights (K0+1 padded)

template <int C1, int C2, int C3>
struct SaSmem {
  static constexpr int W1_BYTES = C2 * C1 * 2;
  static constexpr int W2_BYTES = C3 * C2 * 2;
  static constexpr int H1_BYTES = SA_ROWS * C1 * 2;
  static constexpr int H2_BYTES = SA_ROWS * C2 * 2;
  static constexpr int OFF_W1 = 0;
  static constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
  static constexpr int OFF_H1 = OFF_W2 + W2_BYTES;
  static constexpr int OFF_H2 = OFF_H1 + H1_BYTES;
  static constexpr int OFF_B1 = OFF_H2 + H2_BYTES;          // C2 floats
  static constexpr int OFF_W0 = OFF_B1 + C2 * 4;            // C1*(SA_MAX_K0+1) floats (inline mode)
  static constexpr int TOTAL_PROJ = OFF_W0;
  static constexpr int TOTAL_INLINE = OFF_W0 + C1 * SA_W0_STRIDE * 4;
  static constexpr int TMEM_COLS_USED = C2 + (C3 / 128) * SA_ROWS;
  static constexpr int TMEM_COLS = TMEM_COLS_USED <= 32 ? 32 : TMEM_COLS_USED <= 64 ? 64
                                   : TMEM_COLS_USED <= 128 ? 128 : TMEM_COLS_USED <= 256 ? 256 : 512;
};

// copy a row-major bf16 matrix [rows x K] from global into the blocked K-major smem layout
__device__ __forceinline__ void load_weights_blocked(uint8_t *dst, const __nv_bfloat16 *src, int rows,
                                                     int K, int tid) {
  const int chunks = K / 8;
  for (int e = tid; e < rows * chunks; e += SA_THREADS) {
    const int r = e / chunks, kc = e - r * chunks;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)r * K + kc * 8));
    *reinterpret_cast<uint4 *>(dst + (size_t)kc * rows * 16 + r * 16) = v;
  }
}

// CTAs per SM the kernel is compiled for: two when TMEM columns and shared memory allow it (SA1
// widths), so that one CTA's gather / epilogue overlaps the other's MMAs
template <int C1, int C2, int C3>
constexpr int sa_min_blocks() {
  using L = SaSmem<C1, C2, C3>;
  return (L::TMEM_COLS <= 256 && L::TOTAL_INLINE <= 100 * 1024) ? 2 : 1;
}

template <int C1, int C2, int C3, int NS, bool MODE_PROJ>
__global__ void __launch_bounds__(SA_THREADS, sa_min_blocks<C1, C2, C3>()) sa_fused_kernel(const SaFusedParams p) {
  using L = SaSmem<C1, C2, C3>;
  static_assert(C1 % 16 == 0 && C2 % 16 == 0 && C2 % 64 == 0 && C3 % 128 == 0, "bad widths");
  static_assert(SA_ROWS % NS == 0 && (NS == 16 || NS == 32 || NS == 64), "bad nsample");
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3;          // TMEM lane quarter this warp may access
  const int grp = warp >> 2;       // 0 or 1: which half of the column work this warp takes
  uint8_t *sW1 = smem + L::OFF_W1, *sW2 = smem + L::OFF_W2, *sH1 = smem + L::OFF_H1,
          *sH2 = smem + L::OFF_H2;
  float *sB1 = reinterpret_cast<float *>(smem + L::OFF_B1);
  float *sW0 = reinterpret_cast<float *>(smem + L::OFF_W0);   // [C1][K0+1] (last = b0), inline mode
  const int K0 = 3 + p.Cf;

  // ---- one-time setup ---------------------------------------------------------------------------
  if (warp == 0) tmem_alloc(&tmem_base_smem, L::TMEM_COLS);
  if (tid == 32) {
    mbarrier_init(&mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  load_weights_blocked(sW1, p.W1, C2, C1, tid);
  load_weights_blocked(sW2, p.W2, C3, C2, tid);
  for (int e = tid; e < C2; e += SA_THREADS) sB1[e] = __ldg(p.b1 + e);
  if (!MODE_PROJ) {
    // row c = [w(c,0..K0-1), b0(c), 0...]: the bias rides along as the weight of a constant-1 input
    for (int e = tid; e < C1 * SA_W0_STRIDE; e += SA_THREADS) {
      const int c = e / SA_W0_STRIDE, k = e - c * SA_W0_STRIDE;
      sW0[e] = k < K0 ? __ldg(p.W0 + (size_t)c * K0 + k) : (k == K0 ? __ldg(p.b0 + c) : 0.f);
    }
  }
  fence_proxy_async_smem();      // weights were written through the generic proxy; UMMA reads via async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_d1 = tmem_base;             // [128 lanes x C2 cols]
  const uint32_t tmem_d2 = tmem_base + C2;        // C3/128 blocks of [128 lanes x 128 cols]

  constexpr uint32_t IDESC1 = make_idesc_bf16(128, C2);
  constexpr uint32_t IDESC2 = make_idesc_bf16(128, SA_ROWS);
  const uint32_t aH1 = s2u(sH1), aH2 = s2u(sH2), aW1 = s2u(sW1), aW2 = s2u(sW2);
  unsigned phase = 0;
  const int rows_per_scene = p.np * p.ns;

  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    // ================= stage 0: gather + layer 0 -> H1 (bf16, blocked K-major) ===================
    {
      const int r = q * 32 + lane;                       // row of the tile owned by this lane
      const long long R = (long long)tile * SA_ROWS + r; // global row
      const int b = (int)(R / rows_per_scene);
      const int rr = (int)(R - (long long)b * rows_per_scene);
      const int j = rr / p.ns;                           // centre
      const int i = __ldg(p.idx + R);                    // neighbour point
      constexpr int KC = C1 / 8;                         // 16-byte chunks per row
      constexpr int KC_HALF = KC / 2;
      const int kc0 = grp * KC_HALF;
      if (MODE_PROJ) {
        const float4 *g = reinterpret_cast<const float4 *>(p.G + ((size_t)b * p.n + i) * C1);
        const float4 *h = reinterpret_cast<const float4 *>(p.Hc + ((size_t)b * p.np + j) * C1);
#pragma unroll 4
        for (int kc = kc0; kc < kc0 + KC_HALF; ++kc) {
          const float4 g0 = __ldg(g + 2 * kc), g1 = __ldg(g + 2 * kc + 1);
          const float4 h0 = __ldg(h + 2 * kc), h1 = __ldg(h + 2 * kc + 1);
          uint4 o;
          o.x = pack_bf16x2(fmaxf(g0.x + h0.x, 0.f), fmaxf(g0.y + h0.y, 0.f));
          o.y = pack_bf16x2(fmaxf(g0.z + h0.z, 0.f), fmaxf(g0.w + h0.w, 0.f));
          o.z = pack_bf16x2(fmaxf(g1.x + h1.x, 0.f), fmaxf(g1.y + h1.y, 0.f));
          o.w = pack_bf16x2(fmaxf(g1.z + h1.z, 0.f), fmaxf(g1.w + h1.w, 0.f));
          *reinterpret_cast<uint4 *>(sH1 + (size_t)kc * SA_ROWS * 16 + r * 16) = o;
        }
      } else {
        float in[SA_W0_STRIDE];
        const float *pp = p.xyz + ((size_t)b * p.n + i) * 3;
        const float *cc = p.new_xyz + ((size_t)b * p.np + j) * 3;
        // (p - c) / r, exactly as grouped_xyz -= new_xyz; grouped_xyz /= radius
        in[0] = __fdiv_rn(__ldg(pp + 0) - __ldg(cc + 0), p.radius);
        in[1] = __fdiv_rn(__ldg(pp + 1) - __ldg(cc + 1), p.radius);
        in[2] = __fdiv_rn(__ldg(pp + 2) - __ldg(cc + 2), p.radius);
#pragma unroll
        for (int f = 3; f < SA_W0_STRIDE; ++f)
          in[f] = f < K0 ? __ldg(p.feat + ((size_t)b * p.Cf + (f - 3)) * p.n + i) : (f == K0 ? 1.f : 0.f);
        const int kq = (K0 + 1 + 3) >> 2;                // float4 groups actually used (1..5)
        for (int kc = kc0; kc < kc0 + KC_HALF; ++kc) {
          float acc[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 *w = reinterpret_cast<const float4 *>(sW0 + (size_t)(kc * 8 + c) * SA_W0_STRIDE);
            float a = 0.f;
#pragma unroll
            for (int g4 = 0; g4 < SA_W0_STRIDE / 4; ++g4) {
              if (g4 < kq) {                             // warp-uniform
                const float4 wv = w[g4];
                a = fmaf(wv.x, in[4 * g4 + 0], a);
                a = fmaf(wv.y, in[4 * g4 + 1], a);
                a = fmaf(wv.z, in[4 * g4 + 2], a);
                a = fmaf(wv.w, in[4 * g4 + 3], a);
              }
            }
            acc[c] = fmaxf(a, 0.f);
          }
          uint4 o;
          o.x = pack_bf16x2(acc[0], acc[1]);
          o.y = pack_bf16x2(acc[2], acc[3]);
          o.z = pack_bf16x2(acc[4], acc[5]);
          o.w = pack_bf16x2(acc[6], acc[7]);
          *reinterpret_cast<uint4 *>(sH1 + (size_t)kc * SA_ROWS * 16 + r * 16) = o;
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // ================= stage 1: D1 = H1 . W1'^T  (M=128 rows, N=C2, K=C1) =========================
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < C1 / 16; ++k) {
        const uint64_t da = make_smem_desc(aH1 + k * 2 * SA_ROWS * 16, SA_ROWS * 16, 128, p.desc_swap);
        const uint64_t db = make_smem_desc(aW1 + k * 2 * C2 * 16, C2 * 16, 128, p.desc_swap);
        umma_bf16(tmem_d1, da, db, IDESC1, k > 0);
      }
      umma_commit(&mma_bar);
    }
    mbarrier_wait(&mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ================= epilogue 1: h2 = relu(D1 + b1) -> H2 (thread = row) ========================
    {
      const int r = q * 32 + lane;
      constexpr int COLS_PER_GRP = C2 / 2;
#pragma unroll
      for (int cb = 0; cb < COLS_PER_GRP; cb += 32) {
        const int col0 = grp * COLS_PER_GRP + cb;
        float v[32];
        tmem_ld32(tmem_d1 + ((uint32_t)(q * 32) << 16) + col0, v);
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const float4 ba = *reinterpret_cast<const float4 *>(sB1 + col0 + c8 * 8);
          const float4 bb = *reinterpret_cast<const float4 *>(sB1 + col0 + c8 * 8 + 4);
          uint4 o;
          o.x = pack_bf16x2(fmaxf(v[c8 * 8 + 0] + ba.x, 0.f), fmaxf(v[c8 * 8 + 1] + ba.y, 0.f));
          o.y = pack_bf16x2(fmaxf(v[c8 * 8 + 2] + ba.z, 0.f), fmaxf(v[c8 * 8 + 3] + ba.w, 0.f));
          o.z = pack_bf16x2(fmaxf(v[c8 * 8 + 4] + bb.x, 0.f), fmaxf(v[c8 * 8 + 5] + bb.y, 0.f));
          o.w = pack_bf16x2(fmaxf(v[c8 * 8 + 6] + bb.z, 0.f), fmaxf(v[c8 * 8 + 7] + bb.w, 0.f));
          const int kc = (col0 >> 3) + c8;
          *reinterpret_cast<uint4 *>(sH2 + (size_t)kc * SA_ROWS * 16 + r * 16) = o;
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // ================= stage 2: D2[h] = W2'[h] . H2^T  (M=128 channels, N=128 rows, K=C2) =========
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < C3 / 128; ++h) {
#pragma unroll
        for (int k = 0; k < C2 / 16; ++k) {
          const uint64_t da = make_smem_desc(aW2 + h * 128 * 16 + k * 2 * C3 * 16, C3 * 16, 128, p.desc_swap);
          const uint64_t db = make_smem_desc(aH2 + k * 2 * SA_ROWS * 16, SA_ROWS * 16, 128, p.desc_swap);
          umma_bf16(tmem_d2 + h * SA_ROWS, da, db, IDESC2, k > 0);
        }
      }
      umma_commit(&mma_bar);
    }
    mbarrier_wait(&mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ================= epilogue 2: max over nsample, + b2, ReLU -> out (thread = channel) =========
    {
      constexpr int ITEMS = (C3 / 128) * 2;              // (channel block h, column half)
      const long long R0 = (long long)tile * SA_ROWS;
      const int b = (int)(R0 / rows_per_scene);
      const int j0 = (int)((R0 - (long long)b * rows_per_scene) / NS);   // first centre of the tile
#pragma unroll
      for (int item = grp; item < ITEMS; item += 2) {
        const int h = item >> 1, half = item & 1;
        const int ch = h * 128 + q * 32 + lane;
        const float bias = __ldg(p.b2 + ch);
        float *o = p.out + ((size_t)b * C3 + ch) * p.np + j0;
        if (NS <= 32) {
#pragma unroll
          for (int cb = 0; cb < 64; cb += 32) {
            float v[32];
            tmem_ld32(tmem_d2 + ((uint32_t)(q * 32) << 16) + h * SA_ROWS + half * 64 + cb, v);
#pragma unroll
            for (int gI = 0; gI < 32 / NS; ++gI) {
              float m = v[gI * NS];
#pragma unroll
              for (int t = 1; t < NS; ++t) m = fmaxf(m, v[gI * NS + t]);
              o[(half * 64 + cb) / NS + gI] = fmaxf(m + bias, 0.f);
            }
          }
        } else {  // NS == 64: one centre per 64-column half
          float m = -INFINITY;
#pragma unroll
          for (int cb = 0; cb < 64; cb += 32) {
            float v[32];
            tmem_ld32(tmem_d2 + ((uint32_t)(q * 32) << 16) + h * SA_ROWS + half * 64 + cb, v);
#pragma unroll
            for (int t = 0; t < 32; ++t) m = fmaxf(m, v[t]);
          }
          o[half] = fmaxf(m + bias, 0.f);
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // TMEM / H1 / H2 free for the next tile
    tc_fence_after();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, L::TMEM_COLS);
}

template <int C1, int C2, int C3, int NS, bool MODE_PROJ>
static int launch_sa(const SaFusedParams &p, cudaStream_t stream) {
  using L = SaSmem<C1, C2, C3>;
  auto kern = sa_fused_kernel<C1, C2, C3, NS, MODE_PROJ>;
  const int smem = MODE_PROJ ? L::TOTAL_PROJ : L::TOTAL_INLINE;
  SPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  SPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared));
  // CTAs per SM are bounded by TMEM columns (512 per SM) and shared memory
  int per_sm = 512 / L::TMEM_COLS;
  const int by_smem = (227 * 1024) / (smem + 2048);
  if (per_sm > by_smem) per_sm = by_smem;
  if (per_sm > sa_min_blocks<C1, C2, C3>()) per_sm = sa_min_blocks<C1, C2, C3>();
  if (per_sm < 1) per_sm = 1;
  int grid = kNumSMs * per_sm;
  if (grid > p.num_tiles) grid = p.num_tiles;
  kern<<<grid, SA_THREADS, smem, stream>>>(p);
  SPC_LAUNCH_CHECK("sa_fused_kernel");
  return SPC_OK;
}

}  // namespace spc

using namespace spc;

extern "C" int spc_sa_fused_forward(const float *xyz, const float *new_xyz, const int32_t *idx,
                                    const float *G, const float *Hc, const float *feat,
                                    const float *W0, const float *b0, int Cf, float radius,
                                    const void *W1_bf16, const float *b1, const void *W2_bf16,
                                    const float *b2, int B, int n, int npoint, int nsample, int C1,
                                    int C2, int C3, float *out, void *stream_) {
  SPC_CHECK_ARG(B >= 0 && n >= 1 && npoint >= 0 && nsample >= 1, "sa_fused: bad sizes");
  if (B == 0 || npoint == 0) return SPC_OK;
  SPC_CHECK_ARG(xyz && new_xyz && idx && W1_bf16 && b1 && W2_bf16 && b2 && out, "sa_fused: null pointer");
  const bool proj = G != nullptr;
  SPC_CHECK_ARG(proj ? (Hc != nullptr) : (W0 && b0 && (feat || Cf == 0)), "sa_fused: missing layer-0 operands");
  const long long rows = (long long)B * npoint * nsample;
  if (rows % SA_ROWS != 0 || ((long long)npoint * nsample) % SA_ROWS != 0) {
    set_error("sa_fused: npoint*nsample=%lld is not a multiple of %d", (long long)npoint * nsample, SA_ROWS);
    return SPC_ERR_UNSUPPORTED;
  }
  if (!proj && (Cf < 0 || Cf > SA_MAX_K0 - 3)) {
    set_error("sa_fused: inline mode supports at most %d raw feature channels (got %d)", SA_MAX_K0 - 3, Cf);
    return SPC_ERR_UNSUPPORTED;
  }
  SaFusedParams p;
  p.xyz = xyz; p.new_xyz = new_xyz; p.idx = idx; p.G = G; p.Hc = Hc; p.feat = feat; p.W0 = W0;
  p.b0 = b0; p.Cf = Cf; p.radius = radius;
  p.W1 = (const __nv_bfloat16 *)W1_bf16; p.b1 = b1; p.W2 = (const __nv_bfloat16 *)W2_bf16; p.b2 = b2;
  p.out = out; p.B = B; p.n = n; p.np = npoint; p.ns = nsample;
  p.num_tiles = (int)(rows / SA_ROWS);
  p.desc_swap = 0;
  if (const char *e = getenv("SPC_SA_DESC_SWAP")) p.desc_swap = atoi(e);
  cudaStream_t stream = (cudaStream_t)stream_;
#define SA_TRY(c1, c2, c3, ns)                                                        \
  if (C1 == c1 && C2 == c2 && C3 == c3 && nsample == ns)                              \
    return proj ? launch_sa<c1, c2, c3, ns, true>(p, stream) : launch_sa<c1, c2, c3, ns, false>(p, stream);
  SA_TRY(64, 64, 128, 64)      // SA1
  SA_TRY(64, 64, 128, 32)
  SA_TRY(64, 64, 128, 16)
  SA_TRY(128, 128, 256, 64)
  SA_TRY(128, 128, 256, 32)    // SA2
  SA_TRY(128, 128, 256, 16)    // SA3, SA4
  SA_TRY(128, 128, 128, 64)
  SA_TRY(128, 128, 128, 32)
  SA_TRY(128, 128, 128, 16)    // vote aggregation
#undef SA_TRY
  set_error("sa_fused: no kernel for widths (%d,%d,%d) nsample=%d", C1, C2, C3, nsample);
  return SPC_ERR_UNSUPPORTED;
}
