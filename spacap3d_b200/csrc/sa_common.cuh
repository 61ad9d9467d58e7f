// sa_common.cuh -- PTX wrappers (mbarrier, tcgen05 alloc / mma / commit / ld / fences), the UMMA descriptors of the
// 128-byte-swizzle K-major shared-memory layout and the launch parameters shared by the two fused set-abstraction
// kernels: sa_fused.cu (projected layer 0, SA2-SA4 / vote aggregation) and sa_inline.cu (in-line layer 0, SA1).
#pragma once
#include <cuda_fp16.h>

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
// same, for the gather / epilogue warps, which share their SM sub-partitions with each other: try_wait with a
// suspend-time hint parks the thread in hardware until the phase completes or the hint expires -- ~4 attempts per
// wait in ncu, against a tight spin that stole 25-30 % of the issue slots (round 1) and a nanosleep back-off that
// still spent ~15 % of the kernel's instructions on polling.
__device__ __forceinline__ void mbarrier_wait_relaxed(uint64_t *bar, unsigned parity) {
  // the retry loop stays inside the asm block: try_wait + one predicated branch per attempt (exported through selp
  // into a C++ loop every attempt cost ~8 instructions: selp, setp, divergence bookkeeping)
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SA_RWAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@!p bra SA_RWAIT;\n"
      "}\n" ::"r"(s2u(bar)),
      "r"(parity), "r"(20000u)
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
  const __half *G;        // projected form: (B,n,C1) per-point feature projection (BN scale folded)
  const float *feat;      // in-line form: (B,Cf,n) raw features or nullptr
  const float *W0;        // (C1, 3+Cf) folded, fp32; projected form: Cf = 0, i.e. the xyz columns only
  const float *b0;        // (C1)
  int Cf;
  float radius;           // divide relative xyz by this (1.0 when normalize_xyz is off)
  const __half *W1;       // (C2,C1) folded, fp16 row-major
  const float *b1;        // (C2)
  const __half *W2;       // (C3,C2)
  const float *b2;        // (C3)
  float *out;             // (B,C3,np)
  __half *out_pm;         // optional (B,np,C3): the same result point-major in fp16 (next layer's GEMM input)
  int B, n, np, ns;
  int num_tiles;          // B*np*ns/128
  int min_tiles;          // host-side launch hint (see spc_sa_fused_forward_ex), unused on the device
  unsigned tps_magic;     // ceil(2^32 / tiles_per_scene) when tile / tiles_per_scene == umulhi(tile, magic) for every
                          // tile of this launch, else 0 (the kernels then divide)
};

constexpr int SA_ROWS = 128;        // rows (centre,neighbour pairs) per tile
constexpr int SA_MAX_K0 = 3 + 13;   // the in-line form supports up to 13 raw feature channels (four K steps of 4 inputs)

// scene of a tile (tiles never straddle scenes: npoint*nsample % 128 == 0 is checked at launch)
__device__ __forceinline__ int sa_tile_scene(const SaFusedParams &p, int tile, int tiles_per_scene) {
  return p.tps_magic ? (int)__umulhi((unsigned)tile, p.tps_magic) : tile / tiles_per_scene;
}

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

// max of N (a multiple of 4) registers with four independent chains (a single running maximum is a serial chain of
// N dependent FMNMX in the one warp per SM sub-partition that drains a TMEM block)
template <int N>
__device__ __forceinline__ float sa_max_tree(const float *v) {
  float m0 = v[0], m1 = v[1], m2 = v[2], m3 = v[3];
#pragma unroll
  for (int t = 4; t < N; t += 4) {
    m0 = fmaxf(m0, v[t]); m1 = fmaxf(m1, v[t + 1]); m2 = fmaxf(m2, v[t + 2]); m3 = fmaxf(m3, v[t + 3]);
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// EPILOGUE 2, one 128-channel block: D2 (TMEM lane = output channel, 128 columns = the rows of the tile) -> max over
// the nsample columns of each centre, + b2, relu -> out (B,C3,np) fp32 and, optionally, the point-major fp16 copy.
// The 128/NS results of a lane are consecutive floats of one out row: ONE 8/16-byte store per lane instead of a
// 4-byte store per centre (lane = channel, so every store instruction touches 32 different lines: the scalar form
// cost 32 LSU wavefronts per centre -- with 4-8 centres per tile the largest single user of the L1 data pipe in the
// wide layers, ncu l1tex__data_pipe_lsu_wavefronts).
template <int C3, int NS>
__device__ __forceinline__ void sa_pool_block(const SaFusedParams &p, uint32_t taddr, int b, int j0, int ch) {
  constexpr int NC = SA_ROWS / NS;                   // centres per tile: 2, 4 or 8
  const float bias = __ldg(p.b2 + ch);
  float res[NC];
  float m64 = -INFINITY;
#pragma unroll
  for (int cb = 0; cb < SA_ROWS; cb += 32) {
    float v[32];
    tmem_ld32(taddr + cb, v);
    if (NS <= 32) {
#pragma unroll
      for (int gI = 0; gI < 32 / NS; ++gI) res[cb / NS + gI] = fmaxf(sa_max_tree<NS>(v + gI * NS) + bias, 0.f);
    } else {                                         // NS == 64: two 32-column loads per centre
      m64 = fmaxf(m64, sa_max_tree<32>(v));
      if ((cb & 32) != 0) {
        res[cb / 64] = fmaxf(m64 + bias, 0.f);
        m64 = -INFINITY;
      }
    }
  }
  // j0 and np are multiples of NC (npoint * nsample % 128 == 0) and the tensor base is 16-byte aligned (checked at
  // launch): the NC floats start on an NC*4-byte boundary
  float *o = p.out + ((size_t)b * C3 + ch) * p.np + j0;
  if (NC == 2) {
    *reinterpret_cast<float2 *>(o) = make_float2(res[0], res[1]);
  } else {
#pragma unroll
    for (int c = 0; c < NC; c += 4) *reinterpret_cast<float4 *>(o + c) = make_float4(res[c], res[c + 1], res[c + 2], res[c + 3]);
  }
  if (p.out_pm) {                                    // lanes = consecutive channels => 64-byte coalesced stores
    __half *opm = p.out_pm + ((size_t)b * p.np + j0) * C3 + ch;
#pragma unroll
    for (int c = 0; c < NC; ++c) opm[(size_t)c * C3] = to_f16_sat(res[c]);
  }
}

// grid of a persistent launch: one CTA per resident slot, fewer when the caller asks for a minimum number of tiles
// per CTA.  Every CTA pays a fixed cost (weight staging, TMEM allocation, pipeline fill); with few tiles per CTA
// that cost dominates and a smaller grid spends less SM-time for the same work (slower alone, faster when other
// streams can use the freed SMs).
inline int sa_grid(const SaFusedParams &p, int occ) {
  int grid = kNumSMs * occ;
  if (grid > p.num_tiles) grid = p.num_tiles;
  if (p.min_tiles > 0) grid = max(1, min(grid, (p.num_tiles + p.min_tiles - 1) / p.min_tiles));
  return grid;
}

// in-line form (sa_inline.cu); returns SPC_ERR_UNSUPPORTED (error text set) when no kernel matches the widths
int launch_sa_inline(const SaFusedParams &p, int C1, int C2, int C3, int nsample, cudaStream_t stream);

}  // namespace spc
