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
// same, but yields the issue slot between polls (producer / epilogue warps share their SM
// sub-partitions with each other; a tight spin stole 30 % of the issue slots, ncu)
__device__ __forceinline__ void mbarrier_wait_relaxed(uint64_t *bar, unsigned parity) {
  unsigned done;
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(s2u(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(40);
  }
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

// Same descriptor for the 128-byte-swizzle K-major layout (layout type 2): a row is 128 contiguous
// bytes (64 bf16), 8-row groups are 1024 B apart (SBO), the 16-byte chunk c of row r sits at chunk
// position c ^ (r & 7) (Swizzle<3,4,3>); K beyond 64 elements continues in the next "K atom",
// rows*128 bytes further.  The leading-byte-offset field is unused for swizzled K-major (= 1).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// byte offset of the 16-byte chunk (row, kc) of a [rows x K] bf16 operand in that layout
__device__ __forceinline__ uint32_t sw128_off(int row, int kc, int rows) {
  return (uint32_t)((kc >> 3) * rows * 128 + row * 128 + (((kc & 7) ^ (row & 7)) << 4));
}
// byte offset of K step kk (16 elements) relative to the operand base
__device__ __forceinline__ uint32_t sw128_kstep(int kk, int rows) {
  return (uint32_t)((kk >> 2) * rows * 128 + (kk & 3) * 32);
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

// Gather stage + layer 0 for ONE row of a tile: writes chunks [kc0, kc0+NKC) of row r of H1
// (bf16, blocked K-major).  Shared by the serial and the pipelined kernel.  `i` is the
// neighbour index idx[R] (loaded by the caller so that it can be prefetched a tile ahead).
template <int C1, int NKC>
__device__ __forceinline__ void sa_produce_proj(const SaFusedParams &p, int b, int j, int i, int r, int kc0,
                                                uint8_t *sH1) {
  const float4 *g = reinterpret_cast<const float4 *>(p.G + ((size_t)b * p.n + i) * C1) + 2 * kc0;
  const float4 *h = reinterpret_cast<const float4 *>(p.Hc + ((size_t)b * p.np + j) * C1) + 2 * kc0;
  // all gather loads of this thread are issued before the first use: the stage is bound by L2
  // latency/bandwidth, so memory-level parallelism is what matters
  float4 gv[2 * NKC];
#pragma unroll
  for (int c = 0; c < 2 * NKC; ++c) gv[c] = __ldg(g + c);
#pragma unroll
  for (int kc = 0; kc < NKC; ++kc) {
    const float4 h0 = __ldg(h + 2 * kc), h1 = __ldg(h + 2 * kc + 1);     // per-centre row: L1 hits
    const float4 g0 = gv[2 * kc], g1 = gv[2 * kc + 1];
    uint4 o;
    o.x = pack_bf16x2(fmaxf(g0.x + h0.x, 0.f), fmaxf(g0.y + h0.y, 0.f));
    o.y = pack_bf16x2(fmaxf(g0.z + h0.z, 0.f), fmaxf(g0.w + h0.w, 0.f));
    o.z = pack_bf16x2(fmaxf(g1.x + h1.x, 0.f), fmaxf(g1.y + h1.y, 0.f));
    o.w = pack_bf16x2(fmaxf(g1.z + h1.z, 0.f), fmaxf(g1.w + h1.w, 0.f));
    *reinterpret_cast<uint4 *>(sH1 + (size_t)(kc0 + kc) * SA_ROWS * 16 + r * 16) = o;
  }
}

// in-line layer 0 with KQ float4 groups of inputs: in = [(p-c)/r (3), features (Cf), 1 (bias), 0...]
template <int C1, int NKC, int KQ, bool SW = false>
__device__ __forceinline__ void sa_produce_inline(const SaFusedParams &p, int b, int j, int i, int r, int kc0,
                                                  uint8_t *sH1, const float *sW0, int K0) {
  float in[4 * KQ];
  const float *pp = p.xyz + ((size_t)b * p.n + i) * 3;
  const float *cc = p.new_xyz + ((size_t)b * p.np + j) * 3;
  // (p - c) / r, exactly as grouped_xyz -= new_xyz; grouped_xyz /= radius
  in[0] = __fdiv_rn(__ldg(pp + 0) - __ldg(cc + 0), p.radius);
  in[1] = __fdiv_rn(__ldg(pp + 1) - __ldg(cc + 1), p.radius);
  in[2] = __fdiv_rn(__ldg(pp + 2) - __ldg(cc + 2), p.radius);
  const float *fp = p.feat + (size_t)b * p.Cf * p.n + i;
#pragma unroll
  for (int f = 3; f < 4 * KQ; ++f)
    in[f] = f < K0 ? __ldg(fp + (size_t)(f - 3) * p.n) : (f == K0 ? 1.f : 0.f);
#pragma unroll 2
  for (int kc = kc0; kc < kc0 + NKC; ++kc) {
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 *w = reinterpret_cast<const float4 *>(sW0 + (size_t)(kc * 8 + c) * SA_W0_STRIDE);
      float a = 0.f;
#pragma unroll
      for (int g4 = 0; g4 < KQ; ++g4) {
        const float4 wv = w[g4];
        a = fmaf(wv.x, in[4 * g4 + 0], a);
        a = fmaf(wv.y, in[4 * g4 + 1], a);
        a = fmaf(wv.z, in[4 * g4 + 2], a);
        a = fmaf(wv.w, in[4 * g4 + 3], a);
      }
      acc[c] = fmaxf(a, 0.f);
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]);
    o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]);
    o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4 *>(sH1 + (SW ? sw128_off(r, kc, SA_ROWS) : (uint32_t)(kc * SA_ROWS * 16 + r * 16))) = o;
  }
}

// Projected layer 0 for the pipelined kernel: ONE WARP PER ROW so that the gather of a 4*C1-byte
// G row is a single coalesced request (4 L1 wavefronts instead of 32 with lane = row), 16 rows in
// flight per warp; the 8-byte bf16 results go to the 128B-swizzled H1, where a row is one
// contiguous 128-byte line per K atom => conflict-free stores.
template <int C1>
__device__ __forceinline__ void sa_produce_proj_rowwise(const SaFusedParams &p, int tile, int warp, int lane,
                                                        int my_idx, uint8_t *sH1, int rows_per_scene) {
  constexpr int LPR = C1 / 4;                 // lanes per row (float4 each)
  constexpr int RPI = 32 / LPR;               // rows per warp-wide load
  constexpr int NPASS = 16 / RPI;             // a warp owns 16 rows of the tile
  const int c4 = lane % LPR;
  const int sub = lane / LPR;
  const long long R0 = (long long)tile * SA_ROWS + warp * 16;
  const int b = (int)(R0 / rows_per_scene);
  const int j = (int)(R0 - (long long)b * rows_per_scene) / p.ns;      // the 16 rows share one centre (ns >= 16)
  const float *Gb = p.G + (size_t)b * p.n * C1 + 4 * c4;
  float4 g[NPASS];
#pragma unroll
  for (int t = 0; t < NPASS; ++t) {
    const int i = __shfl_sync(0xffffffffu, my_idx, t * RPI + sub);    // lanes 0..15 hold idx of the 16 rows
    g[t] = __ldg(reinterpret_cast<const float4 *>(Gb + (size_t)i * C1));
  }
  const float4 h = __ldg(reinterpret_cast<const float4 *>(p.Hc + ((size_t)b * p.np + j) * C1 + 4 * c4));
#pragma unroll
  for (int t = 0; t < NPASS; ++t) {
    const int r = warp * 16 + t * RPI + sub;
    uint2 o;
    o.x = pack_bf16x2(fmaxf(g[t].x + h.x, 0.f), fmaxf(g[t].y + h.y, 0.f));
    o.y = pack_bf16x2(fmaxf(g[t].z + h.z, 0.f), fmaxf(g[t].w + h.w, 0.f));
    *reinterpret_cast<uint2 *>(sH1 + sw128_off(r, c4 >> 1, SA_ROWS) + (c4 & 1) * 8) = o;
  }
}

template <int C1, bool MODE_PROJ>
__device__ __forceinline__ void sa_produce_rows(const SaFusedParams &p, int tile, int r, int i, int kc0,
                                                uint8_t *sH1, const float *sW0, int K0, int rows_per_scene) {
  constexpr int NKC = C1 / 16;                       // half of the row's 16-byte chunks
  const long long R = (long long)tile * SA_ROWS + r; // global row
  const int b = (int)(R / rows_per_scene);
  const int j = (int)(R - (long long)b * rows_per_scene) / p.ns;   // centre
  if (MODE_PROJ) {
    sa_produce_proj<C1, NKC>(p, b, j, i, r, kc0, sH1);
  } else {
    switch ((K0 + 1 + 3) >> 2) {                     // float4 groups of inputs actually used (warp-uniform)
      case 1: sa_produce_inline<C1, NKC, 1>(p, b, j, i, r, kc0, sH1, sW0, K0); break;
      case 2: sa_produce_inline<C1, NKC, 2>(p, b, j, i, r, kc0, sH1, sW0, K0); break;
      case 3: sa_produce_inline<C1, NKC, 3>(p, b, j, i, r, kc0, sH1, sW0, K0); break;
      case 4: sa_produce_inline<C1, NKC, 4>(p, b, j, i, r, kc0, sH1, sW0, K0); break;
      default: sa_produce_inline<C1, NKC, 5>(p, b, j, i, r, kc0, sH1, sW0, K0); break;
    }
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
    sa_produce_rows<C1, MODE_PROJ>(p, tile, q * 32 + lane, __ldg(p.idx + (long long)tile * SA_ROWS + q * 32 + lane),
                                   grp * (C1 / 16), sH1, sW0, K0, rows_per_scene);
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

// ================================================================================================
// Warp-specialised, double-buffered version of the same computation.
//
// The serial kernel above runs gather -> MMA1 -> epilogue1 -> MMA2 -> epilogue2 back to back in
// every CTA (tensor pipe 4-14 % active, ncu).  Here the five stages of consecutive tiles overlap:
//   warps 0-7   PRODUCERS   gather + layer 0            -> H1[s]      (s = tile parity)
//   warp  16    MMA ISSUER  one thread: MMA1(k) then MMA2(k-1)        (tcgen05, D in TMEM)
//   warps 8-11  EPILOGUE 1  D1[s] -> relu(+b1) -> bf16  -> H2[s]
//   warps 12-15 EPILOGUE 2  D2[u&1] -> max over nsample, +b2, relu -> out   (u = 128-channel unit)
// Hand-offs are mbarriers: "full" barriers are arrived on by the producing threads (after a
// generic->async proxy fence, because UMMA reads shared memory through the async proxy) or by
// tcgen05.commit; "empty" barriers by the consuming threads / by tcgen05.commit of the MMA that
// read the buffer.  TMEM: D1 double-buffered (2*C2 columns) + a 2-deep ring of 128-column D2
// blocks = at most 512 columns.  Shared memory for the widest layer (128,128,256): W1 32K + W2 64K
// + 2*H1 64K + 2*H2 64K = 224 KB.
// ================================================================================================
constexpr int SAP_PROD_WARPS = 8;
constexpr int SAP_THREADS = 17 * 32;

template <int C1, int C2, int C3>
struct SaPipeSmem {
  static constexpr int W1_BYTES = C2 * C1 * 2;
  static constexpr int W2_BYTES = C3 * C2 * 2;
  static constexpr int H1_BYTES = SA_ROWS * C1 * 2;
  static constexpr int H2_BYTES = SA_ROWS * C2 * 2;
  static constexpr int OFF_W1 = 0;
  static constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
  static constexpr int OFF_H1 = OFF_W2 + W2_BYTES;            // 2 stages
  static constexpr int OFF_H2 = OFF_H1 + 2 * H1_BYTES;        // 2 stages
  static constexpr int OFF_B1 = OFF_H2 + 2 * H2_BYTES;        // C2 floats
  static constexpr int OFF_W0 = OFF_B1 + C2 * 4;              // inline mode only
  static constexpr int TOTAL_PROJ = OFF_W0;
  static constexpr int TOTAL_INLINE = OFF_W0 + C1 * SA_W0_STRIDE * 4;
  static constexpr int TMEM_D2 = 2 * C2;                      // first column of the D2 ring
  static constexpr int TMEM_COLS = 512;
};

__device__ __forceinline__ void mbarrier_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(bar)) : "memory");
}

template <int C1, int C2, int C3, int NS, bool MODE_PROJ>
__global__ void __launch_bounds__(SAP_THREADS, 1) sa_fused_pipe_kernel(const SaFusedParams p) {
  using L = SaPipeSmem<C1, C2, C3>;
  constexpr int NB = C3 / 128;                       // 128-channel output blocks per tile
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // the 128B-swizzle atoms must start on 1024-byte boundaries
  uint8_t *smem = smem_raw + ((1024u - (s2u(smem_raw) & 1023u)) & 1023u);
  // barrier groups, 2 stages each
  __shared__ __align__(8) uint64_t h1_full[2], h1_empty[2], d1_full[2], d1_empty[2], h2_full[2], h2_empty[2],
      d2_full[2], d2_empty[2];
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  uint8_t *sW1 = smem + L::OFF_W1, *sW2 = smem + L::OFF_W2, *sH1 = smem + L::OFF_H1, *sH2 = smem + L::OFF_H2;
  float *sB1 = reinterpret_cast<float *>(smem + L::OFF_B1);
  float *sW0 = reinterpret_cast<float *>(smem + L::OFF_W0);
  const int K0 = 3 + p.Cf;

  // ---- one-time setup ---------------------------------------------------------------------------
  if (warp == 8) tmem_alloc(&tmem_base_smem, L::TMEM_COLS);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      mbarrier_init(&h1_full[s], SAP_PROD_WARPS * 32);
      mbarrier_init(&h1_empty[s], 1);
      mbarrier_init(&d1_full[s], 1);
      mbarrier_init(&d1_empty[s], 128);
      mbarrier_init(&h2_full[s], 128);
      mbarrier_init(&h2_empty[s], 1);
      mbarrier_init(&d2_full[s], 1);
      mbarrier_init(&d2_empty[s], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = tid; e < C2 * (C1 / 8); e += SAP_THREADS) {
    const int r = e / (C1 / 8), kc = e - r * (C1 / 8);
    *reinterpret_cast<uint4 *>(sW1 + sw128_off(r, kc, C2)) =
        __ldg(reinterpret_cast<const uint4 *>(p.W1 + (size_t)r * C1 + kc * 8));
  }
  for (int e = tid; e < C3 * (C2 / 8); e += SAP_THREADS) {
    const int r = e / (C2 / 8), kc = e - r * (C2 / 8);
    *reinterpret_cast<uint4 *>(sW2 + sw128_off(r, kc, C3)) =
        __ldg(reinterpret_cast<const uint4 *>(p.W2 + (size_t)r * C2 + kc * 8));
  }
  for (int e = tid; e < C2; e += SAP_THREADS) sB1[e] = __ldg(p.b1 + e);
  if (!MODE_PROJ) {
    for (int e = tid; e < C1 * SA_W0_STRIDE; e += SAP_THREADS) {
      const int c = e / SA_W0_STRIDE, k = e - c * SA_W0_STRIDE;
      sW0[e] = k < K0 ? __ldg(p.W0 + (size_t)c * K0 + k) : (k == K0 ? __ldg(p.b0 + c) : 0.f);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int rows_per_scene = p.np * p.ns;
  const int nt = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

  if (warp < SAP_PROD_WARPS) {
    // =============================== PRODUCERS ===================================================
    if (MODE_PROJ) {
      // warp per row group: lanes 0..15 carry the neighbour indices of the warp's 16 rows
      auto load_idx = [&](int tile) {
        return lane < 16 ? __ldg(p.idx + (long long)tile * SA_ROWS + warp * 16 + lane) : 0;
      };
      int i_next = load_idx((int)blockIdx.x);
      for (int k = 0; k < nt; ++k) {
        const int s = k & 1, n = k >> 1;
        const int tile = (int)blockIdx.x + k * (int)gridDim.x;
        const int my_idx = i_next;
        if (k + 1 < nt) i_next = load_idx(tile + (int)gridDim.x);              // a tile ahead
        mbarrier_wait_relaxed(&h1_empty[s], (unsigned)(n & 1) ^ 1u);           // MMA1(k-2) has consumed H1[s]
        sa_produce_proj_rowwise<C1>(p, tile, warp, lane, my_idx, sH1 + s * L::H1_BYTES, rows_per_scene);
        fence_proxy_async_smem();
        mbarrier_arrive(&h1_full[s]);
      }
    } else {
      const int r = tid & (SA_ROWS - 1);
      const int half = tid >> 7;                                   // which half of the K chunks
      constexpr int NKC = C1 / 16;
      int i_next = __ldg(p.idx + (long long)blockIdx.x * SA_ROWS + r);
      for (int k = 0; k < nt; ++k) {
        const int s = k & 1, n = k >> 1;
        const int tile = (int)blockIdx.x + k * (int)gridDim.x;
        const int i = i_next;
        if (k + 1 < nt) i_next = __ldg(p.idx + (long long)(tile + (int)gridDim.x) * SA_ROWS + r);
        const long long R = (long long)tile * SA_ROWS + r;
        const int b = (int)(R / rows_per_scene);
        const int j = (int)(R - (long long)b * rows_per_scene) / p.ns;
        mbarrier_wait_relaxed(&h1_empty[s], (unsigned)(n & 1) ^ 1u);
        uint8_t *h1 = sH1 + s * L::H1_BYTES;
        switch ((K0 + 1 + 3) >> 2) {
          case 1: sa_produce_inline<C1, NKC, 1, true>(p, b, j, i, r, half * NKC, h1, sW0, K0); break;
          case 2: sa_produce_inline<C1, NKC, 2, true>(p, b, j, i, r, half * NKC, h1, sW0, K0); break;
          case 3: sa_produce_inline<C1, NKC, 3, true>(p, b, j, i, r, half * NKC, h1, sW0, K0); break;
          case 4: sa_produce_inline<C1, NKC, 4, true>(p, b, j, i, r, half * NKC, h1, sW0, K0); break;
          default: sa_produce_inline<C1, NKC, 5, true>(p, b, j, i, r, half * NKC, h1, sW0, K0); break;
        }
        fence_proxy_async_smem();
        mbarrier_arrive(&h1_full[s]);
      }
    }
  } else if (warp == 16) {
    // =============================== MMA ISSUER (one thread) =====================================
    if (lane == 0) {
      constexpr uint32_t IDESC1 = make_idesc_bf16(128, C2);
      constexpr uint32_t IDESC2 = make_idesc_bf16(128, SA_ROWS);
      const uint32_t aH1 = s2u(sH1), aH2 = s2u(sH2), aW1 = s2u(sW1), aW2 = s2u(sW2);
      for (int k = 0; k <= nt; ++k) {
        if (k < nt) {                                            // D1[s] = H1[s] . W1'^T
          const int s = k & 1, n = k >> 1;
          mbarrier_wait(&h1_full[s], (unsigned)(n & 1));
          mbarrier_wait(&d1_empty[s], (unsigned)(n & 1) ^ 1u);   // epilogue 1 has drained D1[s]
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < C1 / 16; ++kk) {
            const uint64_t da = make_smem_desc_sw128(aH1 + s * L::H1_BYTES + sw128_kstep(kk, SA_ROWS));
            const uint64_t db = make_smem_desc_sw128(aW1 + sw128_kstep(kk, C2));
            umma_bf16(tmem_base + s * C2, da, db, IDESC1, kk > 0);
          }
          umma_commit(&d1_full[s]);
          umma_commit(&h1_empty[s]);
        }
        if (k >= 1) {                                            // D2[u] = W2'[h] . H2[s]^T for tile k-1
          const int kt = k - 1, s = kt & 1, n = kt >> 1;
          mbarrier_wait(&h2_full[s], (unsigned)(n & 1));
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < NB; ++h) {
            const int u = kt * NB + h, st = u & 1, nu = u >> 1;
            mbarrier_wait(&d2_empty[st], (unsigned)(nu & 1) ^ 1u);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < C2 / 16; ++kk) {
              const uint64_t da = make_smem_desc_sw128(aW2 + h * 128 * 128 + sw128_kstep(kk, C3));
              const uint64_t db = make_smem_desc_sw128(aH2 + s * L::H2_BYTES + sw128_kstep(kk, SA_ROWS));
              umma_bf16(tmem_base + L::TMEM_D2 + st * SA_ROWS, da, db, IDESC2, kk > 0);
            }
            umma_commit(&d2_full[st]);
          }
          umma_commit(&h2_empty[s]);
        }
      }
    }
  } else if (warp < 12) {
    // =============================== EPILOGUE 1: D1 -> H2 ========================================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    for (int k = 0; k < nt; ++k) {
      const int s = k & 1, n = k >> 1;
      mbarrier_wait_relaxed(&d1_full[s], (unsigned)(n & 1));
      mbarrier_wait_relaxed(&h2_empty[s], (unsigned)(n & 1) ^ 1u);   // MMA2(k-2) has consumed H2[s]
      tc_fence_after();
      uint8_t *h2 = sH2 + s * L::H2_BYTES;
#pragma unroll
      for (int col0 = 0; col0 < C2; col0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + s * C2 + col0, v);
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const float4 ba = *reinterpret_cast<const float4 *>(sB1 + col0 + c8 * 8);
          const float4 bb = *reinterpret_cast<const float4 *>(sB1 + col0 + c8 * 8 + 4);
          uint4 o;
          o.x = pack_bf16x2(fmaxf(v[c8 * 8 + 0] + ba.x, 0.f), fmaxf(v[c8 * 8 + 1] + ba.y, 0.f));
          o.y = pack_bf16x2(fmaxf(v[c8 * 8 + 2] + ba.z, 0.f), fmaxf(v[c8 * 8 + 3] + ba.w, 0.f));
          o.z = pack_bf16x2(fmaxf(v[c8 * 8 + 4] + bb.x, 0.f), fmaxf(v[c8 * 8 + 5] + bb.y, 0.f));
          o.w = pack_bf16x2(fmaxf(v[c8 * 8 + 6] + bb.z, 0.f), fmaxf(v[c8 * 8 + 7] + bb.w, 0.f));
          *reinterpret_cast<uint4 *>(h2 + sw128_off(r, (col0 >> 3) + c8, SA_ROWS)) = o;
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbarrier_arrive(&h2_full[s]);
      mbarrier_arrive(&d1_empty[s]);
    }
  } else {
    // =============================== EPILOGUE 2: D2 -> max-pool -> out ==========================
    const int q = warp & 3;
    for (int k = 0; k < nt; ++k) {
      const long long R0 = (long long)((int)blockIdx.x + k * (int)gridDim.x) * SA_ROWS;
      const int b = (int)(R0 / rows_per_scene);
      const int j0 = (int)((R0 - (long long)b * rows_per_scene) / NS);   // first centre of the tile
#pragma unroll
      for (int h = 0; h < NB; ++h) {
        const int u = k * NB + h, st = u & 1, nu = u >> 1;
        mbarrier_wait_relaxed(&d2_full[st], (unsigned)(nu & 1));
        tc_fence_after();
        const int ch = h * 128 + q * 32 + lane;
        const float bias = __ldg(p.b2 + ch);
        float *o = p.out + ((size_t)b * C3 + ch) * p.np + j0;
        float m64 = -INFINITY;
#pragma unroll
        for (int cb = 0; cb < SA_ROWS; cb += 32) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + L::TMEM_D2 + st * SA_ROWS + cb, v);
          if (NS <= 32) {
#pragma unroll
            for (int gI = 0; gI < 32 / NS; ++gI) {
              float m = v[gI * NS];
#pragma unroll
              for (int t = 1; t < NS; ++t) m = fmaxf(m, v[gI * NS + t]);
              o[cb / NS + gI] = fmaxf(m + bias, 0.f);
            }
          } else {                                               // NS == 64: two 32-column loads per centre
#pragma unroll
            for (int t = 0; t < 32; ++t) m64 = fmaxf(m64, v[t]);
            if ((cb & 32) != 0) { o[cb / 64] = fmaxf(m64 + bias, 0.f); m64 = -INFINITY; }
          }
        }
        tc_fence_before();
        mbarrier_arrive(&d2_empty[st]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

template <int C1, int C2, int C3, int NS, bool MODE_PROJ>
static int launch_sa_pipe(const SaFusedParams &p, cudaStream_t stream) {
  using L = SaPipeSmem<C1, C2, C3>;
  auto kern = sa_fused_pipe_kernel<C1, C2, C3, NS, MODE_PROJ>;
  const int smem = (MODE_PROJ ? L::TOTAL_PROJ : L::TOTAL_INLINE) + 1024;   // + slack for 1024-byte alignment
  SPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int grid = kNumSMs;
  if (grid > p.num_tiles) grid = p.num_tiles;
  kern<<<grid, SAP_THREADS, smem, stream>>>(p);
  SPC_LAUNCH_CHECK("sa_fused_pipe_kernel");
  return SPC_OK;
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
  // SPC_SA_PIPE=0 selects the serial kernel (kept for A/B measurements)
  const char *pe = getenv("SPC_SA_PIPE");
  const bool pipe = !(pe && atoi(pe) == 0);
#define SA_TRY(c1, c2, c3, ns)                                                                       \
  if (C1 == c1 && C2 == c2 && C3 == c3 && nsample == ns) {                                           \
    if (pipe)                                                                                        \
      return proj ? launch_sa_pipe<c1, c2, c3, ns, true>(p, stream)                                  \
                  : launch_sa_pipe<c1, c2, c3, ns, false>(p, stream);                                \
    return proj ? launch_sa<c1, c2, c3, ns, true>(p, stream) : launch_sa<c1, c2, c3, ns, false>(p, stream); \
  }
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
