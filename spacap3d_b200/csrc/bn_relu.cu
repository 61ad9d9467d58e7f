// bn_relu.cu -- training-mode BatchNorm + ReLU over (B, C, S) fp32 activations, forward and backward (sm_100a).
//
// Replaces, for the shared-MLP blocks in TRAINING mode, the reference's
//     nn.BatchNorm2d (pytorch_utils.py:39-64, cuDNN bn_fw_tr / bn_bw kernels) followed by nn.ReLU(inplace=True)
//     (pytorch_utils.py:11-36) on the (B, C, npoint, nsample) tensors produced by the 1x1 convolutions.
// Profiled on B200 (batch 4 x 40k points, forward+backward): cuDNN's bn_bw_1C11_kernel_new is 30 % of the
// step and bn_fw_tr 11 %, both at < 20 % of HBM bandwidth; together with the separate ReLU forward and
// backward passes the block makes ~14 full passes over tensors of up to 134 MB.  Here:
//   forward : one statistics pass (read y) + one normalise+ReLU pass (read y, write z)
//   backward: one reduction pass (read dz, y; the ReLU mask is recomputed from y) + one apply pass
//             (read dz, y; write dy)
// All four are streaming kernels: float4 loads with UNR independent loads in flight per thread, grids sized in
// multiples of the SM count, fp32 per-thread partial sums of SHIFTED data (x - K, K = first element of the
// channel: no cancellation), per-CTA and cross-CTA combination in fp64.
// Semantics = torch.nn.functional.batch_norm(training=True) + relu: biased variance for normalisation,
// running_mean/var updated with `momentum` and the unbiased variance, save_mean / save_invstd for backward,
// gradient of ReLU taken where the OUTPUT is > 0.
#include "common.cuh"

namespace spc {

constexpr int BN_THREADS = 256;
constexpr int BN_MAX_SPLITS = 64;
constexpr int BN_UNR = 4;

__device__ __forceinline__ double block_sum_double(double v, double *s_red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < BN_THREADS / 32 ? s_red[lane] : 0.0;
#pragma unroll
    for (int o = 4; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in thread 0
}

// grid (splits, C): CTA (s, c) reduces positions [p0, p1) of every batch row of channel c.
// MODE 0: sums of (x-K), (x-K)^2.   MODE 1 (backward): sums of dzm, dzm * xhat.
template <int MODE, bool VEC>
__global__ void __launch_bounds__(BN_THREADS) bn_reduce_kernel(const float *__restrict__ y, const float *__restrict__ dz,
                                                                const float *__restrict__ gamma,
                                                                const float *__restrict__ beta,
                                                                const float *__restrict__ mean,
                                                                const float *__restrict__ invstd, int B, int C, int S,
                                                                int chunk, double *__restrict__ partial) {
  __shared__ double s_red[BN_THREADS / 32];
  const int c = blockIdx.y, s = blockIdx.x, splits = gridDim.x;
  const int p0 = s * chunk, p1 = min(S, p0 + chunk);
  float a0 = 0.f, a1 = 0.f;
  float K = 0.f, mu = 0.f, is = 0.f, g = 0.f, bt = 0.f;
  if (MODE == 0) K = __ldg(y + (size_t)c * S);
  else { mu = __ldg(mean + c); is = __ldg(invstd + c); g = __ldg(gamma + c); bt = __ldg(beta + c); }
  auto acc = [&](float yv, float dv) {
    if (MODE == 0) {
      const float d = yv - K;
      a0 += d;
      a1 = fmaf(d, d, a1);
    } else {
      const float xh = (yv - mu) * is;
      const float z = fmaf(xh, g, bt);
      const float dm = z > 0.f ? dv : 0.f;
      a0 += dm;
      a1 = fmaf(dm, xh, a1);
    }
  };
  for (int b = 0; b < B; ++b) {
    const float *row = y + ((size_t)b * C + c) * S;
    const float *drow = MODE == 1 ? dz + ((size_t)b * C + c) * S : nullptr;
    if (VEC) {
      int p = p0 + threadIdx.x * 4;
      for (; p + (BN_UNR - 1) * BN_THREADS * 4 < p1; p += BN_UNR * BN_THREADS * 4) {
        float4 v[BN_UNR], d[BN_UNR];
#pragma unroll
        for (int u = 0; u < BN_UNR; ++u) {
          v[u] = __ldg(reinterpret_cast<const float4 *>(row + p + u * BN_THREADS * 4));
          if (MODE == 1) d[u] = __ldg(reinterpret_cast<const float4 *>(drow + p + u * BN_THREADS * 4));
        }
#pragma unroll
        for (int u = 0; u < BN_UNR; ++u) {
          acc(v[u].x, MODE == 1 ? d[u].x : 0.f); acc(v[u].y, MODE == 1 ? d[u].y : 0.f);
          acc(v[u].z, MODE == 1 ? d[u].z : 0.f); acc(v[u].w, MODE == 1 ? d[u].w : 0.f);
        }
      }
      for (; p < p1; p += BN_THREADS * 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(row + p));
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == 1) d = __ldg(reinterpret_cast<const float4 *>(drow + p));
        acc(v.x, d.x); acc(v.y, d.y); acc(v.z, d.z); acc(v.w, d.w);
      }
    } else {
      for (int p = p0 + threadIdx.x; p < p1; p += BN_THREADS) acc(__ldg(row + p), MODE == 1 ? __ldg(drow + p) : 0.f);
    }
  }
  const double t0 = block_sum_double((double)a0, s_red);
  const double t1 = block_sum_double((double)a1, s_red);
  if (threadIdx.x == 0) {
    partial[((size_t)c * splits + s) * 2 + 0] = t0;
    partial[((size_t)c * splits + s) * 2 + 1] = t1;
  }
}

__global__ void bn_finalize_forward_kernel(const float *__restrict__ y, const double *__restrict__ partial, int C, int S,
                                           int splits, double M, float eps, float momentum,
                                           float *__restrict__ running_mean, float *__restrict__ running_var,
                                           float *__restrict__ save_mean, float *__restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int s = 0; s < splits; ++s) { s1 += partial[((size_t)c * splits + s) * 2]; s2 += partial[((size_t)c * splits + s) * 2 + 1]; }
  const double K = (double)__ldg(y + (size_t)c * S);
  const double m1 = s1 / M;
  const double mean = K + m1;
  double var = s2 / M - m1 * m1;
  if (var < 0.0) var = 0.0;
  save_mean[c] = (float)mean;
  save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
  if (running_var) {
    const double unbiased = M > 1.0 ? var * M / (M - 1.0) : var;
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

__global__ void bn_finalize_backward_kernel(const double *__restrict__ partial, int C, int splits,
                                            float *__restrict__ dgamma, float *__restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int s = 0; s < splits; ++s) { s1 += partial[((size_t)c * splits + s) * 2]; s2 += partial[((size_t)c * splits + s) * 2 + 1]; }
  dbeta[c] = (float)s1;
  dgamma[c] = (float)s2;
}

// grid (position tiles, B*C).  MODE 0: z = relu(bn(y)).  MODE 1: dy from dz, y and the channel sums.
template <int MODE, bool VEC>
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const float *__restrict__ y, const float *__restrict__ dz,
                                                               const float *__restrict__ gamma, const float *__restrict__ beta,
                                                               const float *__restrict__ mean, const float *__restrict__ invstd,
                                                               const float *__restrict__ dgamma, const float *__restrict__ dbeta,
                                                               int C, int S, float inv_M, float *__restrict__ out) {
  const int rowi = blockIdx.y;
  const int c = rowi % C;
  const float mu = __ldg(mean + c), is = __ldg(invstd + c), g = __ldg(gamma + c), bt = __ldg(beta + c);
  float k0 = 0.f, k1 = 0.f, gs = 0.f;
  if (MODE == 1) { k0 = __ldg(dbeta + c) * inv_M; k1 = __ldg(dgamma + c) * inv_M; gs = g * is; }
  const float *row = y + (size_t)rowi * S;
  const float *drow = MODE == 1 ? dz + (size_t)rowi * S : nullptr;
  float *orow = out + (size_t)rowi * S;
  auto f = [&](float yv, float dv) -> float {
    const float xh = (yv - mu) * is;
    const float z = fmaf(xh, g, bt);
    if (MODE == 0) return fmaxf(z, 0.f);
    const float dm = z > 0.f ? dv : 0.f;
    return gs * ((dm - k0) - xh * k1);
  };
  const int tile = BN_THREADS * 4 * BN_UNR;
  if (VEC) {
    const int base = blockIdx.x * tile + threadIdx.x * 4;
    float4 v[BN_UNR], d[BN_UNR];
#pragma unroll
    for (int u = 0; u < BN_UNR; ++u) {
      const int p = base + u * BN_THREADS * 4;
      if (p < S) {
        v[u] = __ldg(reinterpret_cast<const float4 *>(row + p));
        if (MODE == 1) d[u] = __ldg(reinterpret_cast<const float4 *>(drow + p));
      }
    }
#pragma unroll
    for (int u = 0; u < BN_UNR; ++u) {
      const int p = base + u * BN_THREADS * 4;
      if (p < S) {
        float4 o;
        o.x = f(v[u].x, MODE == 1 ? d[u].x : 0.f); o.y = f(v[u].y, MODE == 1 ? d[u].y : 0.f);
        o.z = f(v[u].z, MODE == 1 ? d[u].z : 0.f); o.w = f(v[u].w, MODE == 1 ? d[u].w : 0.f);
        *reinterpret_cast<float4 *>(orow + p) = o;
      }
    }
  } else {
    for (int p = blockIdx.x * tile + threadIdx.x; p < min(S, (int)(blockIdx.x + 1) * tile); p += BN_THREADS)
      orow[p] = f(__ldg(row + p), MODE == 1 ? __ldg(drow + p) : 0.f);
  }
}

struct BnPlan {
  int splits, chunk;
  bool vec;
};

static BnPlan plan_bn(const void *a, const void *b, const void *c, int C, int S) {
  BnPlan p;
  p.vec = (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
  int splits = ceil_div(4 * kNumSMs, C);                       // ~4 CTAs per SM in total
  splits = max(1, min(min(splits, BN_MAX_SPLITS), ceil_div(S, BN_THREADS * 4)));
  int chunk = ceil_div(S, splits);
  chunk = (chunk + 3) & ~3;
  p.chunk = chunk;
  p.splits = ceil_div(S, chunk);
  return p;
}

// ----------------------------------------------------------------------------------------------------
// Last block of a set-abstraction MLP: BatchNorm + ReLU + max over the nsample neighbours
// (pointnet2_modules.py:256-259: F.max_pool2d(kernel=[1, nsample])).  The (B,C,npoint,nsample) activation is
// never written: the forward emits the pooled (B,C,npoint) tensor, the winning slot and the pre-BN value at
// that slot; the backward needs one pass (read y, write dy) because the incoming gradient is non-zero at one
// slot per group only, so the channel sums come from (B,C,npoint) data.
// A group of NS contiguous floats is handled by NS/4 adjacent lanes (one float4 each): fully coalesced.
// Ties keep the lowest slot, like ATen's max_pool2d (strict '>'), which matters because ball-query padding
// repeats the first neighbour.
// ----------------------------------------------------------------------------------------------------
constexpr int BP_UNR = 4;

template <int NS>
__global__ void __launch_bounds__(BN_THREADS) bn_relu_pool_fwd_kernel(const float *__restrict__ y,
                                                                       const float *__restrict__ gamma,
                                                                       const float *__restrict__ beta,
                                                                       const float *__restrict__ mean,
                                                                       const float *__restrict__ invstd, int C, int np,
                                                                       float *__restrict__ pooled,
                                                                       uint8_t *__restrict__ argmax,
                                                                       float *__restrict__ ymax) {
  constexpr int LPG = NS / 4;                 // lanes per group
  constexpr int GPW = 32 / LPG;               // groups per warp per step
  const int rowi = blockIdx.y, c = rowi % C;
  const float mu = __ldg(mean + c), is = __ldg(invstd + c), g = __ldg(gamma + c), bt = __ldg(beta + c);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPG, gl = lane / LPG;
  const int j0 = (blockIdx.x * (BN_THREADS / 32) + warp) * GPW * BP_UNR + gl;
  const float *row = y + (size_t)rowi * np * NS;
  float4 v[BP_UNR];
#pragma unroll
  for (int u = 0; u < BP_UNR; ++u) {
    const int j = j0 + u * GPW;
    if (j < np) v[u] = __ldg(reinterpret_cast<const float4 *>(row + (size_t)j * NS + sub * 4));
  }
#pragma unroll
  for (int u = 0; u < BP_UNR; ++u) {
    const int j = j0 + u * GPW;             // uniform across the LPG lanes of a group
    const bool ok = j < np;
    const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
    float bz = -INFINITY, by = 0.f;
    int bi = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float z = ok ? fmaf((e[q] - mu) * is, g, bt) : 0.f;
      if (z > bz) { bz = z; bi = sub * 4 + q; by = e[q]; }
    }
#pragma unroll
    for (int o = 1; o < LPG; o <<= 1) {
      const float oz = __shfl_xor_sync(0xffffffffu, bz, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const float oy = __shfl_xor_sync(0xffffffffu, by, o);
      if (oz > bz || (oz == bz && oi < bi)) { bz = oz; bi = oi; by = oy; }
    }
    if (ok && sub == 0) {
      const size_t o = (size_t)rowi * np + j;
      pooled[o] = fmaxf(bz, 0.f);
      argmax[o] = (uint8_t)bi;
      ymax[o] = by;
    }
  }
}

// one CTA per channel: dbeta = sum dm, dgamma = sum dm * xhat over the (B, npoint) winners
__global__ void __launch_bounds__(BN_THREADS) bn_pool_bwd_reduce_kernel(const float *__restrict__ dpool,
                                                                         const float *__restrict__ ymax,
                                                                         const float *__restrict__ gamma,
                                                                         const float *__restrict__ beta,
                                                                         const float *__restrict__ mean,
                                                                         const float *__restrict__ invstd, int B, int C,
                                                                         int np, float *__restrict__ dgamma,
                                                                         float *__restrict__ dbeta) {
  __shared__ double s_red[BN_THREADS / 32];
  const int c = blockIdx.x;
  const float mu = __ldg(mean + c), is = __ldg(invstd + c), g = __ldg(gamma + c), bt = __ldg(beta + c);
  float a0 = 0.f, a1 = 0.f;
  for (int b = 0; b < B; ++b) {
    const size_t base = ((size_t)b * C + c) * np;
    for (int j = threadIdx.x; j < np; j += BN_THREADS) {
      const float xh = (__ldg(ymax + base + j) - mu) * is;
      const float dm = fmaf(xh, g, bt) > 0.f ? __ldg(dpool + base + j) : 0.f;
      a0 += dm;
      a1 = fmaf(dm, xh, a1);
    }
  }
  const double t0 = block_sum_double((double)a0, s_red);
  const double t1 = block_sum_double((double)a1, s_red);
  if (threadIdx.x == 0) { dbeta[c] = (float)t0; dgamma[c] = (float)t1; }
}

template <int NS>
__global__ void __launch_bounds__(BN_THREADS) bn_pool_bwd_apply_kernel(
    const float *__restrict__ y, const float *__restrict__ dpool, const uint8_t *__restrict__ argmax,
    const float *__restrict__ ymax, const float *__restrict__ gamma, const float *__restrict__ beta,
    const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ dgamma,
    const float *__restrict__ dbeta, int C, int np, float inv_M, float *__restrict__ dy) {
  constexpr int LPG = NS / 4;
  constexpr int GPW = 32 / LPG;
  const int rowi = blockIdx.y, c = rowi % C;
  const float mu = __ldg(mean + c), is = __ldg(invstd + c), g = __ldg(gamma + c), bt = __ldg(beta + c);
  const float k0 = __ldg(dbeta + c) * inv_M, k1 = __ldg(dgamma + c) * inv_M, gs = g * is;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPG, gl = lane / LPG;
  const int j0 = (blockIdx.x * (BN_THREADS / 32) + warp) * GPW * BP_UNR + gl;
  const float *row = y + (size_t)rowi * np * NS;
  float *orow = dy + (size_t)rowi * np * NS;
  float4 v[BP_UNR];
  float dp[BP_UNR], ym[BP_UNR];
  int am[BP_UNR];
#pragma unroll
  for (int u = 0; u < BP_UNR; ++u) {
    const int j = j0 + u * GPW;
    if (j < np) {
      v[u] = __ldg(reinterpret_cast<const float4 *>(row + (size_t)j * NS + sub * 4));
      const size_t o = (size_t)rowi * np + j;
      dp[u] = __ldg(dpool + o);
      ym[u] = __ldg(ymax + o);
      am[u] = (int)__ldg(argmax + o);
    }
  }
#pragma unroll
  for (int u = 0; u < BP_UNR; ++u) {
    const int j = j0 + u * GPW;
    if (j >= np) continue;
    const float dwin = fmaf((ym[u] - mu) * is, g, bt) > 0.f ? dp[u] : 0.f;
    const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
    float o4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float xh = (e[q] - mu) * is;
      const float dm = (sub * 4 + q == am[u]) ? dwin : 0.f;
      o4[q] = gs * ((dm - k0) - xh * k1);
    }
    *reinterpret_cast<float4 *>(orow + (size_t)j * NS + sub * 4) = make_float4(o4[0], o4[1], o4[2], o4[3]);
  }
}

}  // namespace spc

using namespace spc;

extern "C" size_t spc_bn_relu_workspace_bytes(int C) {
  return C > 0 ? (size_t)C * BN_MAX_SPLITS * 2 * sizeof(double) : 0;
}

extern "C" int spc_bn_relu_train_forward(const float *y, const float *gamma, const float *beta, int B, int C, int S,
                                         float eps, float momentum, float *running_mean, float *running_var,
                                         float *z, float *save_mean, float *save_invstd, void *workspace,
                                         size_t workspace_bytes, void *stream_) {
  SPC_CHECK_ARG(B >= 1 && C >= 1 && S >= 1, "bn_relu_forward: bad sizes B=%d C=%d S=%d", B, C, S);
  SPC_CHECK_ARG(y && gamma && beta && z && save_mean && save_invstd, "bn_relu_forward: null pointer");
  SPC_CHECK_ARG(workspace && workspace_bytes >= spc_bn_relu_workspace_bytes(C) && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0,
                "bn_relu_forward: workspace of spc_bn_relu_workspace_bytes(C) bytes required");
  SPC_CHECK_ARG((long long)B * C <= 65535 && C <= 65535, "bn_relu_forward: B*C too large");
  cudaStream_t stream = (cudaStream_t)stream_;
  double *partial = reinterpret_cast<double *>(workspace);
  const BnPlan p = plan_bn(y, z, y, C, S);
  dim3 rgrid(p.splits, C);
  if (p.vec) bn_reduce_kernel<0, true><<<rgrid, BN_THREADS, 0, stream>>>(y, nullptr, gamma, beta, nullptr, nullptr, B, C, S, p.chunk, partial);
  else bn_reduce_kernel<0, false><<<rgrid, BN_THREADS, 0, stream>>>(y, nullptr, gamma, beta, nullptr, nullptr, B, C, S, p.chunk, partial);
  bn_finalize_forward_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(y, partial, C, S, p.splits, (double)B * S, eps, momentum,
                                                                   running_mean, running_var, save_mean, save_invstd);
  dim3 agrid(ceil_div(S, BN_THREADS * 4 * BN_UNR), B * C);
  if (p.vec) bn_apply_kernel<0, true><<<agrid, BN_THREADS, 0, stream>>>(y, nullptr, gamma, beta, save_mean, save_invstd, nullptr, nullptr, C, S, 0.f, z);
  else bn_apply_kernel<0, false><<<agrid, BN_THREADS, 0, stream>>>(y, nullptr, gamma, beta, save_mean, save_invstd, nullptr, nullptr, C, S, 0.f, z);
  SPC_LAUNCH_CHECK("bn_relu_train_forward");
  return SPC_OK;
}

extern "C" int spc_bn_relu_train_backward(const float *dz, const float *y, const float *gamma, const float *beta,
                                          const float *save_mean, const float *save_invstd, int B, int C, int S,
                                          float *dy, float *dgamma, float *dbeta, void *workspace,
                                          size_t workspace_bytes, void *stream_) {
  SPC_CHECK_ARG(B >= 1 && C >= 1 && S >= 1, "bn_relu_backward: bad sizes B=%d C=%d S=%d", B, C, S);
  SPC_CHECK_ARG(dz && y && gamma && beta && save_mean && save_invstd && dy && dgamma && dbeta, "bn_relu_backward: null pointer");
  SPC_CHECK_ARG(workspace && workspace_bytes >= spc_bn_relu_workspace_bytes(C) && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0,
                "bn_relu_backward: workspace of spc_bn_relu_workspace_bytes(C) bytes required");
  SPC_CHECK_ARG((long long)B * C <= 65535 && C <= 65535, "bn_relu_backward: B*C too large");
  cudaStream_t stream = (cudaStream_t)stream_;
  double *partial = reinterpret_cast<double *>(workspace);
  const BnPlan p = plan_bn(y, dz, dy, C, S);
  dim3 rgrid(p.splits, C);
  if (p.vec) bn_reduce_kernel<1, true><<<rgrid, BN_THREADS, 0, stream>>>(y, dz, gamma, beta, save_mean, save_invstd, B, C, S, p.chunk, partial);
  else bn_reduce_kernel<1, false><<<rgrid, BN_THREADS, 0, stream>>>(y, dz, gamma, beta, save_mean, save_invstd, B, C, S, p.chunk, partial);
  bn_finalize_backward_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(partial, C, p.splits, dgamma, dbeta);
  dim3 agrid(ceil_div(S, BN_THREADS * 4 * BN_UNR), B * C);
  const float inv_M = (float)(1.0 / ((double)B * S));
  if (p.vec) bn_apply_kernel<1, true><<<agrid, BN_THREADS, 0, stream>>>(y, dz, gamma, beta, save_mean, save_invstd, dgamma, dbeta, C, S, inv_M, dy);
  else bn_apply_kernel<1, false><<<agrid, BN_THREADS, 0, stream>>>(y, dz, gamma, beta, save_mean, save_invstd, dgamma, dbeta, C, S, inv_M, dy);
  SPC_LAUNCH_CHECK("bn_relu_train_backward");
  return SPC_OK;
}

// ---- BatchNorm + ReLU + max over nsample (last block of a set-abstraction MLP) ---------------------------
static bool pool_ns_supported(int ns) { return ns == 16 || ns == 32 || ns == 64; }

extern "C" int spc_bn_relu_maxpool_train_forward(const float *y, const float *gamma, const float *beta, int B, int C,
                                                 int npoint, int nsample, float eps, float momentum,
                                                 float *running_mean, float *running_var, float *pooled,
                                                 uint8_t *argmax, float *ymax, float *save_mean, float *save_invstd,
                                                 void *workspace, size_t workspace_bytes, void *stream_) {
  SPC_CHECK_ARG(B >= 1 && C >= 1 && npoint >= 1 && nsample >= 1, "bn_relu_maxpool_forward: bad sizes");
  if (!pool_ns_supported(nsample) || (reinterpret_cast<uintptr_t>(y) & 15)) {
    set_error("bn_relu_maxpool: nsample=%d not in {16,32,64} or y not 16-byte aligned", nsample);
    return SPC_ERR_UNSUPPORTED;
  }
  SPC_CHECK_ARG(y && gamma && beta && pooled && argmax && ymax && save_mean && save_invstd, "bn_relu_maxpool_forward: null pointer");
  SPC_CHECK_ARG(workspace && workspace_bytes >= spc_bn_relu_workspace_bytes(C) && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0,
                "bn_relu_maxpool_forward: workspace of spc_bn_relu_workspace_bytes(C) bytes required");
  SPC_CHECK_ARG((long long)B * C <= 65535 && (long long)npoint * nsample < (1LL << 31), "bn_relu_maxpool_forward: too large");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int S = npoint * nsample;
  double *partial = reinterpret_cast<double *>(workspace);
  const BnPlan p = plan_bn(y, y, y, C, S);
  bn_reduce_kernel<0, true><<<dim3(p.splits, C), BN_THREADS, 0, stream>>>(y, nullptr, gamma, beta, nullptr, nullptr, B, C, S, p.chunk, partial);
  bn_finalize_forward_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(y, partial, C, S, p.splits, (double)B * S, eps, momentum,
                                                                   running_mean, running_var, save_mean, save_invstd);
#define BP_FWD(NS)                                                                                          \
  bn_relu_pool_fwd_kernel<NS><<<dim3(ceil_div(npoint, (BN_THREADS / 32) * (32 / (NS / 4)) * BP_UNR), B * C), \
                                BN_THREADS, 0, stream>>>(y, gamma, beta, save_mean, save_invstd, C, npoint, pooled, argmax, ymax)
  if (nsample == 16) BP_FWD(16); else if (nsample == 32) BP_FWD(32); else BP_FWD(64);
#undef BP_FWD
  SPC_LAUNCH_CHECK("bn_relu_maxpool_train_forward");
  return SPC_OK;
}

extern "C" int spc_bn_relu_maxpool_train_backward(const float *dpool, const uint8_t *argmax, const float *ymax,
                                                  const float *y, const float *gamma, const float *beta,
                                                  const float *save_mean, const float *save_invstd, int B, int C,
                                                  int npoint, int nsample, float *dy, float *dgamma, float *dbeta,
                                                  void *stream_) {
  SPC_CHECK_ARG(B >= 1 && C >= 1 && npoint >= 1 && nsample >= 1, "bn_relu_maxpool_backward: bad sizes");
  if (!pool_ns_supported(nsample) || ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dy)) & 15)) {
    set_error("bn_relu_maxpool: nsample=%d not in {16,32,64} or tensors not 16-byte aligned", nsample);
    return SPC_ERR_UNSUPPORTED;
  }
  SPC_CHECK_ARG(dpool && argmax && ymax && y && gamma && beta && save_mean && save_invstd && dy && dgamma && dbeta,
                "bn_relu_maxpool_backward: null pointer");
  SPC_CHECK_ARG((long long)B * C <= 65535 && C <= 65535, "bn_relu_maxpool_backward: too large");
  cudaStream_t stream = (cudaStream_t)stream_;
  bn_pool_bwd_reduce_kernel<<<C, BN_THREADS, 0, stream>>>(dpool, ymax, gamma, beta, save_mean, save_invstd, B, C, npoint, dgamma, dbeta);
  const float inv_M = (float)(1.0 / ((double)B * npoint * nsample));
#define BP_BWD(NS)                                                                                            \
  bn_pool_bwd_apply_kernel<NS><<<dim3(ceil_div(npoint, (BN_THREADS / 32) * (32 / (NS / 4)) * BP_UNR), B * C),  \
                                 BN_THREADS, 0, stream>>>(y, dpool, argmax, ymax, gamma, beta, save_mean, save_invstd, \
                                                          dgamma, dbeta, C, npoint, inv_M, dy)
  if (nsample == 16) BP_BWD(16); else if (nsample == 32) BP_BWD(32); else BP_BWD(64);
#undef BP_BWD
  SPC_LAUNCH_CHECK("bn_relu_maxpool_train_backward");
  return SPC_OK;
}
