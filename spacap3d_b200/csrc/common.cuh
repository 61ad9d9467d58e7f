// common.cuh -- shared helpers for libspacap3d_ops.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/spacap3d_ops.h"

#ifndef __CUDA_ARCH_LIST__
#define __CUDA_ARCH_LIST__ 1000
#endif

namespace spc {

constexpr int kNumSMs = 148;  // B200

// thread-local error text (spc_last_error)
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define SPC_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      spc::set_error(__VA_ARGS__);      \
      return SPC_ERR_INVALID_ARG;       \
    }                                   \
  } while (0)

#define SPC_CUDA(expr)                                   \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return spc::cuda_fail(_e, #expr); \
  } while (0)

#define SPC_LAUNCH_CHECK(name)                                      \
  do {                                                              \
    cudaError_t _e = cudaGetLastError();                            \
    if (_e != cudaSuccess) return spc::cuda_fail(_e, name);         \
  } while (0)

// Squared distance in the reference's exact rounding order (SASS of
// ball_query_gpu.cu:31-32 / sampling_gpu.cu:103-104 / interpolate_gpu.cu:33):
//   t = dy*dy (FMUL);  t = fma(dx,dx,t);  d = fma(dz,dz,t)
// Intrinsics pin the contraction so a compiler change cannot re-associate it.
__device__ __forceinline__ float sqdist_ref(float ax, float ay, float az, float bx, float by,
                                            float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// opt_n_threads of the reference (include/cuda_utils.h:15-19): largest power of two <= n,
// capped at 512 -- it defines the FPS tie-break order, so it is part of the numerical contract.
static inline int ref_opt_n_threads(int work_size) {
  int t = 1;
  while (t * 2 <= work_size && t * 2 <= 512) t *= 2;
  return t;
}

}  // namespace spc
