// fp_vote.cu -- small fused kernels around the point-major (fp16) eval path of the detector's
// feature-propagation stage; the 1x1 convs and the voting tail are csrc/pm_linear.cu.  They replace chains of tiny
// ATen launches (sqrt/add/reciprocal/sum/div, cat, transpose copies, norm/div; ~18 % of the SM time
// of a forward, profiles/r1_sm_cycles_per_kernel_one_forward.csv) with one kernel each.
#include <cuda_fp16.h>

#include "common.cuh"

namespace spc {

// ------------------------------------------------------------------------------------------------
// three_nn + inverse-distance weights in one kernel.
// PointnetFPModule.forward (reference pointnet2_modules.py:398-402):
//   dist, idx = three_nn(unknown, known);  dist_recip = 1.0 / (dist + 1e-8)
//   norm = sum(dist_recip, dim=2);          weight = dist_recip / norm
// Same arithmetic (IEEE sqrt / div, left-to-right sum); idx is bit-identical to three_nn.
// ------------------------------------------------------------------------------------------------
constexpr int NW_THREADS = 256;
constexpr int NW_SPLIT = 8;          // lanes cooperating on one unknown point (power of two <= 32)
constexpr int NW_SPLIT_LOG2 = 3;
constexpr int NW_TILE = 1024;

// insert candidate (d,i) into the ascending triple, ordering by (distance, index): the reference's
// strict '<' scan in ascending k keeps the lower index on ties, which is exactly this order
__device__ __forceinline__ void nn3_insert(float d, int i, float &b1, int &i1, float &b2, int &i2, float &b3,
                                           int &i3) {
  if (d < b1 || (d == b1 && i < i1)) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = i; }
  else if (d < b2 || (d == b2 && i < i2)) { b3 = b2; i3 = i2; b2 = d; i2 = i; }
  else if (d < b3 || (d == b3 && i < i3)) { b3 = d; i3 = i; }
}

__global__ void __launch_bounds__(NW_THREADS) three_nn_weights_kernel(const float *__restrict__ unknown,
                                                                       const float *__restrict__ known, int n,
                                                                       int m, int32_t *__restrict__ idx,
                                                                       float *__restrict__ weight) {
  __shared__ float sx[NW_TILE], sy[NW_TILE], sz[NW_TILE];
  const int b = blockIdx.y;
  const int sub = threadIdx.x & (NW_SPLIT - 1);
  const int j = blockIdx.x * (NW_THREADS / NW_SPLIT) + (threadIdx.x >> NW_SPLIT_LOG2);
  const float *U = unknown + (size_t)b * n * 3;
  const float *K = known + (size_t)b * m * 3;
  const bool ok = j < n;
  const float ux = ok ? __ldg(U + 3 * j + 0) : 0.f, uy = ok ? __ldg(U + 3 * j + 1) : 0.f,
              uz = ok ? __ldg(U + 3 * j + 2) : 0.f;
  float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
  int i1 = 0x7fffffff, i2 = 0x7fffffff, i3 = 0x7fffffff;
  for (int base = 0; base < m; base += NW_TILE) {
    const int tile = min(NW_TILE, m - base);
    __syncthreads();
    for (int e = threadIdx.x; e < tile * 3; e += NW_THREADS) {
      const float v = __ldg(K + (size_t)base * 3 + e);
      const int pt = e / 3, comp = e - pt * 3;
      (comp == 0 ? sx : comp == 1 ? sy : sz)[pt] = v;
    }
    __syncthreads();
    for (int k = sub; k < tile; k += NW_SPLIT) {       // each lane of the group scans every NW_SPLIT-th point
      const float d = sqdist_ref(ux, uy, uz, sx[k], sy[k], sz[k]);
      if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = base + k; }
      else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = base + k; }
      else if (d < b3) { b3 = d; i3 = base + k; }
    }
  }
  // merge the four partial triples (butterfly over the quad), ordered by (distance, index)
#pragma unroll
  for (int o = 1; o < NW_SPLIT; o <<= 1) {
    const float c1 = __shfl_xor_sync(0xffffffffu, b1, o), c2 = __shfl_xor_sync(0xffffffffu, b2, o),
                c3 = __shfl_xor_sync(0xffffffffu, b3, o);
    const int k1 = __shfl_xor_sync(0xffffffffu, i1, o), k2 = __shfl_xor_sync(0xffffffffu, i2, o),
              k3 = __shfl_xor_sync(0xffffffffu, i3, o);
    nn3_insert(c1, k1, b1, i1, b2, i2, b3, i3);
    nn3_insert(c2, k2, b1, i1, b2, i2, b3, i3);
    nn3_insert(c3, k3, b1, i1, b2, i2, b3, i3);
  }
  if (ok && sub == 0) {
    const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b1), 1e-8f));
    const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b2), 1e-8f));
    const float r3 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b3), 1e-8f));
    const float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
    float *w = weight + ((size_t)b * n + j) * 3;
    int32_t *oi = idx + ((size_t)b * n + j) * 3;
    w[0] = __fdiv_rn(r1, norm); w[1] = __fdiv_rn(r2, norm); w[2] = __fdiv_rn(r3, norm);
    oi[0] = i1; oi[1] = i2; oi[2] = i3;
  }
}

// ------------------------------------------------------------------------------------------------
// three_interpolate + concat with the skip features, point-major fp16 in and out:
//   X[b, j, 0:C2]      = sum_t w[b,j,t] * known_pm[b, idx[b,j,t], :]     (fp32 accumulate)
//   X[b, j, C2:C2+C1]  = skip_pm[b, j, :]
// (reference: three_interpolate + torch.cat, pointnet2_modules.py:404-416, there on (B,C,n) fp32).
// One warp per point; lanes stride over 8-channel (16-byte) chunks => coalesced rows.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) interp_cat_pm_kernel(const __half *__restrict__ known_pm,
                                                            const int32_t *__restrict__ idx,
                                                            const float *__restrict__ weight,
                                                            const __half *__restrict__ skip_pm, int n,
                                                            int m, int C2, int C1,
                                                            __half *__restrict__ X) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= n) return;
  const int32_t *ix = idx + ((size_t)b * n + j) * 3;
  const float *w = weight + ((size_t)b * n + j) * 3;
  const int a1 = __ldg(ix), a2 = __ldg(ix + 1), a3 = __ldg(ix + 2);
  const float w1 = __ldg(w), w2 = __ldg(w + 1), w3 = __ldg(w + 2);
  const uint4 *r1 = reinterpret_cast<const uint4 *>(known_pm + ((size_t)b * m + a1) * C2);
  const uint4 *r2 = reinterpret_cast<const uint4 *>(known_pm + ((size_t)b * m + a2) * C2);
  const uint4 *r3 = reinterpret_cast<const uint4 *>(known_pm + ((size_t)b * m + a3) * C2);
  uint4 *out = reinterpret_cast<uint4 *>(X + ((size_t)b * n + j) * (C2 + C1));
  for (int c = lane; c < C2 / 8; c += 32) {
    const uint4 v1 = __ldg(r1 + c), v2 = __ldg(r2 + c), v3 = __ldg(r3 + c);
    const uint32_t p1[4] = {v1.x, v1.y, v1.z, v1.w}, p2[4] = {v2.x, v2.y, v2.z, v2.w}, p3[4] = {v3.x, v3.y, v3.z, v3.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      // same contraction order as three_interpolate: fma(p3,w3, fma(p1,w1, p2*w2)); a convex combination of
      // fp16 values cannot leave the fp16 range
      const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&p1[q]));
      const float2 f2 = __half22float2(*reinterpret_cast<const __half2 *>(&p2[q]));
      const float2 f3 = __half22float2(*reinterpret_cast<const __half2 *>(&p3[q]));
      const float lo = fmaf(f3.x, w3, fmaf(f1.x, w1, f2.x * w2));
      const float hi = fmaf(f3.y, w3, fmaf(f1.y, w1, f2.y * w2));
      __half2 pk = __floats2half2_rn(lo, hi);
      o[q] = *reinterpret_cast<uint32_t *>(&pk);
    }
    out[c] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  const uint4 *sk = reinterpret_cast<const uint4 *>(skip_pm + ((size_t)b * n + j) * C1);
  for (int c = lane; c < C1 / 8; c += 32) out[C2 / 8 + c] = __ldg(sk + c);
}

}  // namespace spc

using namespace spc;

extern "C" int spc_three_nn_weights(const float *unknown, const float *known, int B, int n, int m,
                                    int32_t *idx, float *weight, void *stream_) {
  SPC_CHECK_ARG(B >= 0 && n >= 0 && m >= 3, "three_nn_weights: need m >= 3 known points");
  if (B == 0 || n == 0) return SPC_OK;
  SPC_CHECK_ARG(unknown && known && idx && weight, "three_nn_weights: null pointer");
  SPC_CHECK_ARG(B <= 65535, "three_nn_weights: B too large");
  three_nn_weights_kernel<<<dim3(ceil_div(n, NW_THREADS / NW_SPLIT), B), NW_THREADS, 0, (cudaStream_t)stream_>>>(
      unknown, known, n, m, idx, weight);
  SPC_LAUNCH_CHECK("three_nn_weights_kernel");
  return SPC_OK;
}

extern "C" int spc_interp_cat_pm(const void *known_pm_f16, const int32_t *idx, const float *weight,
                                 const void *skip_pm_f16, int B, int n, int m, int C2, int C1,
                                 void *X_f16, void *stream_) {
  SPC_CHECK_ARG(B >= 0 && n >= 0 && m >= 1 && C2 >= 8 && C1 >= 0, "interp_cat_pm: bad sizes");
  SPC_CHECK_ARG(C2 % 8 == 0 && C1 % 8 == 0, "interp_cat_pm: channel counts must be multiples of 8");
  if (B == 0 || n == 0) return SPC_OK;
  SPC_CHECK_ARG(known_pm_f16 && idx && weight && X_f16 && (skip_pm_f16 || C1 == 0), "interp_cat_pm: null pointer");
  SPC_CHECK_ARG(B <= 65535, "interp_cat_pm: B too large");
  interp_cat_pm_kernel<<<dim3(ceil_div(n, 8), B), 256, 0, (cudaStream_t)stream_>>>(
      (const __half *)known_pm_f16, idx, weight, (const __half *)skip_pm_f16, n, m, C2, C1,
      (__half *)X_f16);
  SPC_LAUNCH_CHECK("interp_cat_pm_kernel");
  return SPC_OK;
}
