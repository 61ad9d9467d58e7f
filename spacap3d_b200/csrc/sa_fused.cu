// sa_fused.cu -- fused set-abstraction forward (eval mode): grouping + relative-xyz
// normalisation + shared MLP (3 x [1x1 conv + folded BN + ReLU]) + max-pool over nsample in ONE
// warp-specialised kernel, the two wide 1x1 convs on tcgen05 tensor cores with TMEM accumulators
// (sm_100a).
//
// Replaces, for PointnetSAModuleVotes.forward in eval mode (reference pointnet2_modules.py:244-271):
//   QueryAndGroup's two group_points launches + sub + div + cat (pointnet2_utils.py:351-362),
//   SharedMLP = 3 x (cuDNN conv, cuDNN BN, ReLU) (pytorch_utils.py:11-36) and F.max_pool2d --
// i.e. ~14 kernels that each stream a (B, C, npoint, nsample) activation through HBM
// (~270 MB/scene unfused vs ~8 MB compulsory).
//
// Algorithm (per tile of 128 rows, a row = one (centre, neighbour) pair):
//   layer 0  h1 = relu(W0' . [ (p_i - c_j)/r , f_i ] + b0)         CUDA cores, in the gather stage
//            in-line form (few input channels, SA1): evaluated directly from xyz and raw features.
//            projected form: conv0 is linear and its feature part only depends on the POINT, not on
//            the pair, so it is hoisted out of the grouping: G[i] = W0f' . f_i is one plain GEMM per
//            layer over the n points (npoint*nsample/n = 4..16x fewer MACs) and the kernel evaluates
//            h1 = relu(G[idx] + W0x' . (p_i - c_j)/r + b0) -- the xyz term stays in fp32 inside the
//            kernel (it is a difference of nearby points; G is fp16).
//   layer 1  D1[128 rows x C2]  = H1[128 x C1] . W1'^T             tcgen05.mma, M=128, N=C2
//            h2 = relu(D1 + b1) -> fp16 -> shared memory (thread per row, TMEM lane = row)
//   layer 2  D2[C3 x 128 rows]  = W2'[C3 x C2] . H2^T              tcgen05.mma, TRANSPOSED so that a
//            TMEM lane is an output CHANNEL and the 128 columns are the rows of the tile: the
//            max over the nsample neighbours of a centre is then a register-only reduction.
//   out[b, c, j] = relu(max_k D2[c, j*ns+k] + b2[c])   (ReLU and +b commute with max)
// BN (eval) is folded on the host: W' = diag(gamma/sqrt(var+eps)) W, b = beta - mean*scale.
//
// Execution: a persistent grid of warp-specialised CTAs (8 producer warps, 4 + 4 epilogue warps, 1 MMA-issuing
// warp, mbarrier hand-offs, H1 / H2 / D1 double-buffered).  The narrow in-line configuration (SA1: 94 KB of shared
// memory, one D2 block => 256 TMEM columns, 56 registers) runs TWO CTAs per SM; the wide ones (up to 224 KB) one.
// The in-line layer-0 weights reach the kernel by value in its parameters (constant bank -> uniform registers) when
// the caller supplies host copies (spc_sa_fused_forward_ex), else they are staged in shared memory.
//
// Shared-memory operands: canonical UMMA K-major layout with 128-byte swizzle (a row = 128
// contiguous bytes per 64-element K atom, chunk c of row r at position c ^ (r & 7)).  With it both
// a warp writing one whole row (row-wise gather) and 8 lanes writing the same chunk of 8
// consecutive rows (epilogue, lane = row) are bank-conflict free, and no TMA descriptor is needed
// for gathered data.
#include <cuda_fp16.h>
#include <type_traits>

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
// same, for the producer / epilogue warps, which share their SM sub-partitions with each other: a tight spin stole
// 25-30 % of the issue slots (ncu, round 1) and a nanosleep back-off still spent ~15 % of the kernel's instructions
// on polling (SYNCS + NANOSLEEP + their branches, round 2) -- and the whole pipeline is issue-slot bound.  try_wait
// with a suspend-time hint parks the thread in hardware until the phase completes (it wakes at once) or the hint
// expires: a handful of polls per wait, no wake-up delay.
__device__ __forceinline__ void mbarrier_wait_relaxed(uint64_t *bar, unsigned parity) {
  unsigned done;
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(s2u(bar)), "r"(parity), "r"(20000u)
        : "memory");
    if (done) break;
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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
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

// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): c_format (bits 4-5) = 1 (F32), a_format (7-9) =
// b_format (10-12) = 0 (F16; 1 would be BF16), both operands K-major, N >> 3 at bit 17, M >> 4 at bit 24.
// fp16 operands carry 11 significant bits (bf16: 8): the fused MLP lands ~8x closer to the fp32 reference for the
// same tensor-pipe rate and shared-memory footprint.  Conversions saturate (cvt ... .satfinite) instead of
// producing inf above 65504, a range the BN-folded activations of this network never approach.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout: [0,14) start address >> 4,
// [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4, [46,48) version = 1,
// [61,64) layout type) for the 128-byte-swizzle K-major layout (type 2): a row is 128 contiguous
// bytes (64 fp16), 8-row groups are 1024 B apart (SBO), the 16-byte chunk c of row r sits at chunk
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
// byte offset of the 16-byte chunk (row, kc) of a [rows x K] fp16 operand in that layout
__device__ __forceinline__ uint32_t sw128_off(int row, int kc, int rows) {
  return (uint32_t)((kc >> 3) * rows * 128 + row * 128 + (((kc & 7) ^ (row & 7)) << 4));
}
// byte offset of K step kk (16 elements) relative to the operand base
__device__ __forceinline__ uint32_t sw128_kstep(int kk, int rows) {
  return (uint32_t)((kk >> 2) * rows * 128 + (kk & 3) * 32);
}

// relu + round-to-nearest fp16 conversion (saturating) + packing of two floats in ONE instruction
// (cvt.rn.satfinite.relu.f16x2.f32: first source -> upper half, second source -> lower half)
__device__ __forceinline__ uint32_t pack_relu_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ __half to_f16_sat(float v) {
  uint16_t r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return __ushort_as_half(r);
}

// ---------------------------------------------------------------------------------------------
struct SaFusedParams {
  const float *xyz;       // (B,n,3)
  const float *new_xyz;   // (B,np,3)
  const int32_t *idx;     // (B,np,ns)
  const __half *G; // projected form: (B,n,C1) per-point feature projection (BN scale folded)
  const float *feat;      // in-line form: (B,Cf,n) raw features or nullptr
  const float *W0;        // (C1, 3+Cf) folded, fp32; projected form: Cf = 0, i.e. the xyz columns only
  const float *b0;        // (C1)
  int Cf;
  float radius;           // divide relative xyz by this (1.0 when normalize_xyz is off)
  const __half *W1;  // (C2,C1) folded, fp16 row-major
  const float *b1;          // (C2)
  const __half *W2;  // (C3,C2)
  const float *b2;          // (C3)
  float *out;             // (B,C3,np)
  __half *out_pm;  // optional (B,np,C3): the same result point-major in fp16 (next layer's GEMM input)
  int B, n, np, ns;
  int num_tiles;          // B*np*ns/128
  int min_tiles;          // host-side launch hint (see spc_sa_fused_forward_ex), unused on the device
  // in-line form, optional: the folded layer-0 weights BY VALUE, [c][k] with k = K0 holding the bias.  Kernel
  // parameters live in constant bank 0, so `fmaf(p.w0c[const], x, acc)` compiles to an FFMA with a c[0x0][..]
  // operand: no shared-memory load, no scoreboard wait (the LDS.128 weight loads and the FMAs waiting for them
  // were the largest producer stall in ncu after the gather fix).
  int use_w0c;
  float w0c[768];
};
constexpr int SA_W0C_MAX = 768;

constexpr int SA_ROWS = 128;        // rows (centre,neighbour pairs) per tile
constexpr int SA_MAX_K0 = 3 + 16;   // MODE_INLINE supports up to 16 raw feature channels
constexpr int SA_W0_STRIDE = 20;    // floats per channel row of the inline layer-0 weights (K0+1 padded)

// In-line layer 0 with exactly NIN inputs: in = [(p-c)/r (3), features (Cf), 1 (bias)], NIN = 4+Cf.
// Weights sit in shared memory transposed per 8-channel chunk, sW0t[kc][k][8], so that one input
// updates 8 accumulators from two broadcast 16-byte reads and no FMA is spent on padding.
// The loads of a row are ISSUED a tile ahead (raw values stay in registers) and only FINISHED -- subtraction, scaling --
// when the tile is computed: doing the arithmetic at issue time made every producer thread wait for its own gather
// (ncu: stall_long_sb on the subtraction was as large as the whole layer-0 math).  Tiles never straddle scenes
// (npoint*nsample % 128 == 0 is checked at launch), so scene / centre follow from 32-bit arithmetic on the tile index.
template <int NIN, int NS>
__device__ __forceinline__ void sa_inline_issue(const SaFusedParams &p, int b, int tile_in_scene, int r, int i,
                                                float (&raw)[NIN + 2]) {
  const int j = (tile_in_scene * SA_ROWS + r) / NS;
  const float *pp = p.xyz + ((size_t)b * p.n + i) * 3;
  const float *cc = p.new_xyz + ((size_t)b * p.np + j) * 3;
  raw[0] = __ldg(pp + 0); raw[1] = __ldg(pp + 1); raw[2] = __ldg(pp + 2);
  raw[3] = __ldg(cc + 0); raw[4] = __ldg(cc + 1); raw[5] = __ldg(cc + 2);
  const float *fp = p.feat + (size_t)b * p.Cf * p.n + i;
#pragma unroll
  for (int f = 3; f < NIN - 1; ++f) raw[3 + f] = __ldg(fp + (size_t)(f - 3) * p.n);
}

template <int NIN>
__device__ __forceinline__ void sa_inline_finish(const float (&raw)[NIN + 2], float inv_r, float (&in)[NIN]) {
  // (p - c) / r as a multiplication by 1/r (this path feeds a fp16 MLP: a 1-ulp difference to the
  // reference's true division is far below the rounding of the next step)
  in[0] = (raw[0] - raw[3]) * inv_r;
  in[1] = (raw[1] - raw[4]) * inv_r;
  in[2] = (raw[2] - raw[5]) * inv_r;
#pragma unroll
  for (int f = 3; f < NIN - 1; ++f) in[f] = raw[3 + f];
  in[NIN - 1] = 1.f;                                  // the folded bias rides along as the last "input"
}

template <int C1, int NKC, int NIN>
__device__ __forceinline__ void sa_inline_compute(const float (&in)[NIN], int r, int kc0, uint8_t *sH1,
                                                  const float *sW0t) {
#pragma unroll 2
  for (int kc = kc0; kc < kc0 + NKC; ++kc) {
    const float4 *w = reinterpret_cast<const float4 *>(sW0t + (size_t)kc * SA_W0_STRIDE * 8);
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll
    for (int k = 0; k < NIN; ++k) {
      const float4 wa = w[2 * k], wb = w[2 * k + 1];
      acc[0] = fmaf(wa.x, in[k], acc[0]); acc[1] = fmaf(wa.y, in[k], acc[1]);
      acc[2] = fmaf(wa.z, in[k], acc[2]); acc[3] = fmaf(wa.w, in[k], acc[3]);
      acc[4] = fmaf(wb.x, in[k], acc[4]); acc[5] = fmaf(wb.y, in[k], acc[5]);
      acc[6] = fmaf(wb.z, in[k], acc[6]); acc[7] = fmaf(wb.w, in[k], acc[7]);
    }
    uint4 o;
    o.x = pack_relu_f16x2(acc[0], acc[1]);
    o.y = pack_relu_f16x2(acc[2], acc[3]);
    o.z = pack_relu_f16x2(acc[4], acc[5]);
    o.w = pack_relu_f16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4 *>(sH1 + sw128_off(r, kc, SA_ROWS)) = o;
  }
}

// Same layer with the weights read from the kernel parameters (constant bank); KC0 = first 8-channel chunk.
template <int C1, int NKC, int NIN, int KC0>
__device__ __forceinline__ void sa_inline_compute_const(const SaFusedParams &p, const float (&in)[NIN], int r,
                                                        uint8_t *sH1) {
#pragma unroll
  for (int kc = KC0; kc < KC0 + NKC; ++kc) {
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float a = p.w0c[(kc * 8 + c) * NIN + NIN - 1];                       // folded bias
#pragma unroll
      for (int k = 0; k < NIN - 1; ++k) a = fmaf(p.w0c[(kc * 8 + c) * NIN + k], in[k], a);
      acc[c] = a;
    }
    uint4 o;
    o.x = pack_relu_f16x2(acc[0], acc[1]);
    o.y = pack_relu_f16x2(acc[2], acc[3]);
    o.z = pack_relu_f16x2(acc[4], acc[5]);
    o.w = pack_relu_f16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4 *>(sH1 + sw128_off(r, kc, SA_ROWS)) = o;
  }
}

// Projected layer 0: ONE WARP PER ROW PAIR so that the gather of a 2*C1-byte G row (fp16) is a
// coalesced request (2 L1 wavefronts per row instead of 16+ with lane = row), 16 rows in flight per
// warp.  Lane = 8 consecutive channels of one row: one 16-byte load, 8 x (3 FMA + add + relu) with
// this lane's xyz weights / bias in registers, one 16-byte store into the swizzled H1.
template <int C1, int NS>
__device__ __forceinline__ void sa_produce_proj_rowwise(const SaFusedParams &p, int tile, int warp, int lane,
                                                        int my_idx, const float (&wx)[8][3], const float (&wb)[8],
                                                        uint8_t *sH1, int tiles_per_scene) {
  constexpr int LPR = C1 / 8;                 // lanes per row (8 fp16 = 16 bytes each)
  constexpr int RPI = 32 / LPR;               // rows per warp-wide load
  constexpr int NPASS = 16 / RPI;             // a warp owns 16 rows of the tile
  const int kc = lane % LPR;
  const int sub = lane / LPR;
  const int b = tile / tiles_per_scene;                                // tiles never straddle scenes
  const int j = ((tile - b * tiles_per_scene) * SA_ROWS + warp * 16) / NS;   // the 16 rows share one centre (ns >= 16)
  const __half *Gb = p.G + (size_t)b * p.n * C1 + 8 * kc;
  const float *Pb = p.xyz + (size_t)b * p.n * 3;
  const float *cc = p.new_xyz + ((size_t)b * p.np + j) * 3;
  const float cx = __ldg(cc), cy = __ldg(cc + 1), cz = __ldg(cc + 2);
  const float inv_r = 1.0f / p.radius;
  // two batches of NPASS/2 rows: enough loads in flight to cover the L2 latency without
  // exceeding the 96-register budget of the 17-warp CTA
  constexpr int HB = NPASS / 2 > 0 ? NPASS / 2 : 1;
#pragma unroll
  for (int t0 = 0; t0 < NPASS; t0 += HB) {
    uint4 g[HB];
    float rx[HB], ry[HB], rz[HB];
#pragma unroll
    for (int t = 0; t < HB; ++t) {
      const int i = __shfl_sync(0xffffffffu, my_idx, (t0 + t) * RPI + sub);   // lanes 0..15 hold idx of the 16 rows
      g[t] = __ldg(reinterpret_cast<const uint4 *>(Gb + (size_t)i * C1));
      // (p - c) / r as a multiplication by 1/r: this path is the fp16 one (rtol 1e-2), a 1-ulp
      // difference to the reference's true division is irrelevant here
      rx[t] = (__ldg(Pb + 3 * i + 0) - cx) * inv_r;
      ry[t] = (__ldg(Pb + 3 * i + 1) - cy) * inv_r;
      rz[t] = (__ldg(Pb + 3 * i + 2) - cz) * inv_r;
    }
#pragma unroll
    for (int t = 0; t < HB; ++t) {
      const int r = warp * 16 + (t0 + t) * RPI + sub;
      const uint32_t gw[4] = {g[t].x, g[t].y, g[t].z, g[t].w};
      uint32_t ow[4];
#pragma unroll
      for (int c2 = 0; c2 < 4; ++c2) {
        const float2 g01 = __half22float2(*reinterpret_cast<const __half2 *>(&gw[c2]));                // fp16 -> f32
        const float g0 = g01.x, g1 = g01.y;
        const float v0 = fmaf(wx[2 * c2][2], rz[t], fmaf(wx[2 * c2][1], ry[t], fmaf(wx[2 * c2][0], rx[t], g0 + wb[2 * c2])));
        const float v1 = fmaf(wx[2 * c2 + 1][2], rz[t], fmaf(wx[2 * c2 + 1][1], ry[t], fmaf(wx[2 * c2 + 1][0], rx[t], g1 + wb[2 * c2 + 1])));
        ow[c2] = pack_relu_f16x2(v0, v1);
      }
      *reinterpret_cast<uint4 *>(sH1 + sw128_off(r, kc, SA_ROWS)) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
}

// ================================================================================================
// Warp-specialised, double-buffered kernel.
//
// A first, serial version ran gather -> MMA1 -> epilogue1 -> MMA2 -> epilogue2 back to back in every
// CTA (tensor pipe 4-14 % active, ncu).  Here the five stages of consecutive tiles overlap:
//   warps 0-7   PRODUCERS   gather + layer 0            -> H1[s]      (s = tile parity)
//   warp  16    MMA ISSUER  one thread: MMA1(k) then MMA2(k-1)        (tcgen05, D in TMEM)
//   warps 8-11  EPILOGUE 1  D1[s] -> relu(+b1) -> fp16  -> H2[s]
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

template <int C1, int C2, int C3, int OCC = 1>
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
  // OCC = 2 (two CTAs per SM; only the narrow SA1 widths fit): a single D2 block, so that a CTA needs
  // 2*C2 + 128 <= 256 of the SM's 512 TMEM columns; the other resident CTA covers the lost MMA2 / epilogue-2 overlap
  static constexpr int D2_STAGES = OCC == 2 ? 1 : 2;
  static constexpr int TMEM_COLS = OCC == 2 ? 256 : 512;
};

__device__ __forceinline__ void mbarrier_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(bar)) : "memory");
}

// Row-major fp16 weights (ROWS, 8*KCH) -> the swizzled UMMA layout in shared memory; 8 x 16-byte loads in flight
// per thread.
template <int ROWS, int KCH, int NT>
__device__ __forceinline__ void sa_stage_weights(uint8_t *dst, const __half *src, int t) {
  constexpr int N = ROWS * KCH;
  for (int base = 0; base < N; base += NT * 8) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = base + u * NT + t;
      if (e < N) v[u] = __ldg(reinterpret_cast<const uint4 *>(src) + e);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = base + u * NT + t;
      if (e < N) {
        const int r = e / KCH, kc = e - r * KCH;
        *reinterpret_cast<uint4 *>(dst + sw128_off(r, kc, ROWS)) = v[u];
      }
    }
  }
}

template <int C1, int C2, int C3, int NS, bool MODE_PROJ, int OCC>
__global__ void __launch_bounds__(SAP_THREADS, OCC) sa_fused_pipe_kernel(const SaFusedParams p) {
  using L = SaPipeSmem<C1, C2, C3, OCC>;
  constexpr int D2S = L::D2_STAGES;
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
  // Only the barriers and the TMEM allocation are needed by everybody.  The 24-96 KB of folded weights are staged
  // by the NON-producer warps (all loads of a pass issued before the first store: one L2 round trip per pass
  // instead of one per element) while the producers already gather and compute their first tile; ncu had the
  // serial staging + its barrier at ~20 % of the SA2 kernel's samples and more for the smaller layers.
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < SAP_PROD_WARPS) {
    if (!MODE_PROJ) {
      // sW0t[kc][k][c8]: weight of input k for channel kc*8+c8; k == K0 holds the folded bias
      for (int e = tid; e < C1 * SA_W0_STRIDE; e += SAP_PROD_WARPS * 32) {
        const int kc = e / (SA_W0_STRIDE * 8), rem = e - kc * (SA_W0_STRIDE * 8);
        const int k = rem >> 3, c = kc * 8 + (rem & 7);
        sW0[e] = k < K0 ? __ldg(p.W0 + (size_t)c * K0 + k) : (k == K0 ? __ldg(p.b0 + c) : 0.f);
      }
      asm volatile("bar.sync 2, %0;" ::"n"(SAP_PROD_WARPS * 32) : "memory");
    }
  } else {
    constexpr int NT = SAP_THREADS - SAP_PROD_WARPS * 32;     // 288 staging threads
    const int t = tid - SAP_PROD_WARPS * 32;
    sa_stage_weights<C2, C1 / 8, NT>(sW1, p.W1, t);
    sa_stage_weights<C3, C2 / 8, NT>(sW2, p.W2, t);
    for (int e = t; e < C2; e += NT) sB1[e] = __ldg(p.b1 + e);
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
  }
  const uint32_t tmem_base = tmem_base_smem;
  const int tiles_per_scene = (p.np * NS) / SA_ROWS;
  const int nt = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

  if (warp < SAP_PROD_WARPS) {
    // =============================== PRODUCERS ===================================================
    if (MODE_PROJ) {
      // warp per row group: lanes 0..15 carry the neighbour indices of the warp's 16 rows; every
      // lane keeps the xyz weights and bias of its 8 channels in registers
      constexpr int LPR = C1 / 8;
      float wx[8][3], wb[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int ch = (lane % LPR) * 8 + c;
        wx[c][0] = __ldg(p.W0 + ch * 3 + 0);
        wx[c][1] = __ldg(p.W0 + ch * 3 + 1);
        wx[c][2] = __ldg(p.W0 + ch * 3 + 2);
        wb[c] = __ldg(p.b0 + ch);
      }
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
        sa_produce_proj_rowwise<C1, NS>(p, tile, warp, lane, my_idx, wx, wb, sH1 + s * L::H1_BYTES, tiles_per_scene);
        fence_proxy_async_smem();
        mbarrier_arrive(&h1_full[s]);
      }
    } else {
      // lane = row (two threads per row, half of the channels each); the neighbour index is fetched
      // two tiles ahead and the gathered inputs one tile ahead, so that a tile's math never waits
      // for its own loads
      const int r = tid & (SA_ROWS - 1);
      const int half = tid >> 7;
      constexpr int NKC = C1 / 16;
      const float inv_r = 1.0f / p.radius;
      auto run = [&](auto nin_tag) {
        constexpr int NIN = decltype(nin_tag)::value;
        const int t0 = (int)blockIdx.x, dt = (int)gridDim.x;
        float raw_next[NIN + 2];
        int i1 = __ldg(p.idx + (long long)t0 * SA_ROWS + r);
        // (scene, tile within the scene) of the tile whose loads are issued next, advanced without a division
        int nb = t0 / tiles_per_scene, nts = t0 - nb * tiles_per_scene;
        const int db = dt / tiles_per_scene, dts = dt - db * tiles_per_scene;
        auto advance = [&]() { nb += db; nts += dts; if (nts >= tiles_per_scene) { nts -= tiles_per_scene; ++nb; } };
        sa_inline_issue<NIN, NS>(p, nb, nts, r, i1, raw_next);
        advance();
        i1 = nt > 1 ? __ldg(p.idx + (long long)(t0 + dt) * SA_ROWS + r) : 0;
        for (int k = 0; k < nt; ++k) {
          const int s = k & 1, n = k >> 1;
          const int tile = t0 + k * dt;
          float in[NIN];
          sa_inline_finish<NIN>(raw_next, inv_r, in);
          if (k + 1 < nt) {
            sa_inline_issue<NIN, NS>(p, nb, nts, r, i1, raw_next);
            advance();
            if (k + 2 < nt) i1 = __ldg(p.idx + (long long)(tile + 2 * dt) * SA_ROWS + r);
          }
          mbarrier_wait_relaxed(&h1_empty[s], (unsigned)(n & 1) ^ 1u);
          if (C1 * NIN <= SA_W0C_MAX && p.use_w0c) {
            if (half == 0) sa_inline_compute_const<C1, NKC, NIN, 0>(p, in, r, sH1 + s * L::H1_BYTES);
            else sa_inline_compute_const<C1, NKC, NIN, NKC>(p, in, r, sH1 + s * L::H1_BYTES);
          } else {
            sa_inline_compute<C1, NKC, NIN>(in, r, half * NKC, sH1 + s * L::H1_BYTES, sW0);
          }
          fence_proxy_async_smem();
          mbarrier_arrive(&h1_full[s]);
        }
      };
      switch (K0 + 1) {                                            // exact input count (warp-uniform)
        case 4: run(std::integral_constant<int, 4>{}); break;
        case 5: run(std::integral_constant<int, 5>{}); break;
        case 6: run(std::integral_constant<int, 6>{}); break;
        case 7: run(std::integral_constant<int, 7>{}); break;
        case 8: run(std::integral_constant<int, 8>{}); break;
        case 9: run(std::integral_constant<int, 9>{}); break;
        case 10: run(std::integral_constant<int, 10>{}); break;
        case 11: run(std::integral_constant<int, 11>{}); break;
        case 12: run(std::integral_constant<int, 12>{}); break;
        default:
          if constexpr (OCC == 1) {                                // the two-CTA build (56 registers) only takes <= 8 raw channels
            switch (K0 + 1) {
              case 13: run(std::integral_constant<int, 13>{}); break;
              case 14: run(std::integral_constant<int, 14>{}); break;
              case 15: run(std::integral_constant<int, 15>{}); break;
              case 16: run(std::integral_constant<int, 16>{}); break;
              case 17: run(std::integral_constant<int, 17>{}); break;
              case 18: run(std::integral_constant<int, 18>{}); break;
              case 19: run(std::integral_constant<int, 19>{}); break;
              case 20: run(std::integral_constant<int, 20>{}); break;
              default: break;
            }
          }
          break;
      }
    }
  } else if (warp == 16) {
    // =============================== MMA ISSUER (one thread) =====================================
    if (lane == 0) {
      constexpr uint32_t IDESC1 = make_idesc_f16(128, C2);
      constexpr uint32_t IDESC2 = make_idesc_f16(128, SA_ROWS);
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
            umma_f16(tmem_base + s * C2, da, db, IDESC1, kk > 0);
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
            const int u = kt * NB + h, st = u % D2S, nu = u / D2S;
            mbarrier_wait(&d2_empty[st], (unsigned)(nu & 1) ^ 1u);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < C2 / 16; ++kk) {
              const uint64_t da = make_smem_desc_sw128(aW2 + h * 128 * 128 + sw128_kstep(kk, C3));
              const uint64_t db = make_smem_desc_sw128(aH2 + s * L::H2_BYTES + sw128_kstep(kk, SA_ROWS));
              umma_f16(tmem_base + L::TMEM_D2 + st * SA_ROWS, da, db, IDESC2, kk > 0);
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
          o.x = pack_relu_f16x2(v[c8 * 8 + 0] + ba.x, v[c8 * 8 + 1] + ba.y);
          o.y = pack_relu_f16x2(v[c8 * 8 + 2] + ba.z, v[c8 * 8 + 3] + ba.w);
          o.z = pack_relu_f16x2(v[c8 * 8 + 4] + bb.x, v[c8 * 8 + 5] + bb.y);
          o.w = pack_relu_f16x2(v[c8 * 8 + 6] + bb.z, v[c8 * 8 + 7] + bb.w);
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
      const int tile = (int)blockIdx.x + k * (int)gridDim.x;
      const int b = tile / tiles_per_scene;
      const int j0 = ((tile - b * tiles_per_scene) * SA_ROWS) / NS;      // first centre of the tile
#pragma unroll
      for (int h = 0; h < NB; ++h) {
        const int u = k * NB + h, st = u % D2S, nu = u / D2S;
        mbarrier_wait_relaxed(&d2_full[st], (unsigned)(nu & 1));
        tc_fence_after();
        const int ch = h * 128 + q * 32 + lane;
        const float bias = __ldg(p.b2 + ch);
        float *o = p.out + ((size_t)b * C3 + ch) * p.np + j0;
        // optional point-major fp16 copy: lanes = consecutive channels => 64-byte coalesced stores
        __half *opm = p.out_pm ? p.out_pm + ((size_t)b * p.np + j0) * C3 + ch : nullptr;
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
              const float res = fmaxf(m + bias, 0.f);
              o[cb / NS + gI] = res;
              if (opm) opm[(size_t)(cb / NS + gI) * C3] = to_f16_sat(res);
            }
          } else {                                               // NS == 64: two 32-column loads per centre
#pragma unroll
            for (int t = 0; t < 32; ++t) m64 = fmaxf(m64, v[t]);
            if ((cb & 32) != 0) {
              const float res = fmaxf(m64 + bias, 0.f);
              o[cb / 64] = res;
              if (opm) opm[(size_t)(cb / 64) * C3] = to_f16_sat(res);
              m64 = -INFINITY;
            }
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

template <int C1, int C2, int C3, int NS, bool MODE_PROJ, int OCC>
static int launch_sa_pipe_occ(const SaFusedParams &p, cudaStream_t stream) {
  using L = SaPipeSmem<C1, C2, C3, OCC>;
  auto kern = sa_fused_pipe_kernel<C1, C2, C3, NS, MODE_PROJ, OCC>;
  const int smem = (MODE_PROJ ? L::TOTAL_PROJ : L::TOTAL_INLINE) + 1024;   // + slack for 1024-byte alignment
  SPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int grid = kNumSMs * OCC;
  if (grid > p.num_tiles) grid = p.num_tiles;
  // Throughput knob: every CTA pays a fixed cost (weight staging, TMEM allocation, pipeline fill); with few
  // tiles per CTA that cost dominates and a smaller grid spends less SM-time for the same work (slower alone,
  // faster when other streams can use the freed SMs).
  const int min_tiles = p.min_tiles;
  if (min_tiles > 0) grid = max(1, min(grid, (p.num_tiles + min_tiles - 1) / min_tiles));
  kern<<<grid, SAP_THREADS, smem, stream>>>(p);
  SPC_LAUNCH_CHECK("sa_fused_pipe_kernel");
  return SPC_OK;
}

// The narrow in-line configuration (SA1: 64,64,128) needs 94 KB of shared memory and 256 TMEM columns, so two CTAs
// share an SM when the registers allow it (<= 56 per thread): 34 resident warps instead of 17 hide the latencies
// that keep the single CTA at ~45 % issue utilisation.
template <int C1, int C2, int C3, int NS, bool MODE_PROJ>
static int launch_sa_pipe(const SaFusedParams &p, cudaStream_t stream) {
  if constexpr (!MODE_PROJ && 2 * C2 + 128 <= 256 && 2 * (SaPipeSmem<C1, C2, C3, 2>::TOTAL_INLINE + 1024) <= 227 * 1024) {
    if (p.Cf <= 8) return launch_sa_pipe_occ<C1, C2, C3, NS, MODE_PROJ, 2>(p, stream);
  }
  return launch_sa_pipe_occ<C1, C2, C3, NS, MODE_PROJ, 1>(p, stream);
}

}  // namespace spc

using namespace spc;

extern "C" int spc_sa_fused_forward(const float *xyz, const float *new_xyz, const int32_t *idx,
                                    const void *G_f16, const float *feat, const float *W0,
                                    const float *b0, int Cf, float radius, const void *W1_f16,
                                    const float *b1, const void *W2_f16, const float *b2, int B, int n,
                                    int npoint, int nsample, int C1, int C2, int C3, float *out,
                                    void *out_pm_f16, void *stream_) {
  return spc_sa_fused_forward_ex(xyz, new_xyz, idx, G_f16, feat, W0, b0, nullptr, nullptr, Cf, radius, W1_f16, b1,
                                 W2_f16, b2, B, n, npoint, nsample, C1, C2, C3, out, out_pm_f16, 0, stream_);
}

extern "C" int spc_sa_fused_forward_ex(const float *xyz, const float *new_xyz, const int32_t *idx,
                                       const void *G_f16, const float *feat, const float *W0,
                                       const float *b0, const float *W0_host, const float *b0_host, int Cf,
                                       float radius, const void *W1_f16, const float *b1, const void *W2_f16,
                                       const float *b2, int B, int n, int npoint, int nsample, int C1, int C2,
                                       int C3, float *out, void *out_pm_f16, int min_tiles_per_cta,
                                       void *stream_) {
  SPC_CHECK_ARG(B >= 0 && n >= 1 && npoint >= 0 && nsample >= 1, "sa_fused: bad sizes");
  SPC_CHECK_ARG(min_tiles_per_cta >= 0 && min_tiles_per_cta <= 4096, "sa_fused: min_tiles_per_cta %d out of range",
                min_tiles_per_cta);
  if (B == 0 || npoint == 0) return SPC_OK;
  SPC_CHECK_ARG(xyz && new_xyz && idx && W0 && b0 && W1_f16 && b1 && W2_f16 && b2 && out,
                "sa_fused: null pointer");
  const bool proj = G_f16 != nullptr;
  SPC_CHECK_ARG(proj ? (Cf == 0) : (feat || Cf == 0), "sa_fused: missing layer-0 operands");
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
  p.xyz = xyz; p.new_xyz = new_xyz; p.idx = idx; p.G = (const __half *)G_f16; p.feat = feat;
  p.W0 = W0; p.b0 = b0; p.Cf = Cf; p.radius = radius;
  p.W1 = (const __half *)W1_f16; p.b1 = b1; p.W2 = (const __half *)W2_f16; p.b2 = b2;
  p.out = out; p.out_pm = (__half *)out_pm_f16; p.B = B; p.n = n; p.np = npoint; p.ns = nsample;
  p.num_tiles = (int)(rows / SA_ROWS);
  p.min_tiles = min_tiles_per_cta;
  p.use_w0c = 0;
  if (!proj && W0_host && b0_host && C1 * (4 + Cf) <= SA_W0C_MAX) {
    const int K0 = 3 + Cf, NIN = K0 + 1;
    for (int c = 0; c < C1; ++c) {
      for (int k = 0; k < K0; ++k) p.w0c[c * NIN + k] = W0_host[c * K0 + k];
      p.w0c[c * NIN + K0] = b0_host[c];
    }
    p.use_w0c = 1;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
#define SA_TRY(c1, c2, c3, ns)                                                                       \
  if (C1 == c1 && C2 == c2 && C3 == c3 && nsample == ns)                                             \
    return proj ? launch_sa_pipe<c1, c2, c3, ns, true>(p, stream)                                    \
                : launch_sa_pipe<c1, c2, c3, ns, false>(p, stream);
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
