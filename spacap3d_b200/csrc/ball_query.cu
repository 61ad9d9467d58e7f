// ball_query.cu -- warp-cooperative ordered ball query (sm_100a).
//
// Replaces query_ball_point_kernel (reference ball_query_gpu.cu:9-54): one THREAD per centre,
// one block per scene, every thread streaming all N points from L1/L2.
//
// Here: one WARP owns CPW centres; a CTA stages a tile of the scene's points in shared memory
// (SoA, conflict-free) once for all of its warps.  Each lane tests ONE point against the warp's
// CPW centres per step (distances in the reference's exact FMA order, F3), __ballot_sync gives
// the hit mask in ascending point order, __popc of the lower lanes gives each hit its output
// slot -- so the "first nsample hits in index order" contract holds with no atomics.  A warp
// stops scanning when all of its centres are full (early exit); rows are assembled in shared
// memory and written back coalesced, including the reference's padding (first hit repeated)
// and its all-zero row for an empty ball (F9: the reference gets that from torch::zeros).
#include "common.cuh"

namespace spc {

constexpr int BQ_THREADS = 256;
constexpr int BQ_WARPS = BQ_THREADS / 32;
constexpr int BQ_TILE = 2048;  // points staged per tile: 24 KB

template <int CPW>
__global__ void __launch_bounds__(BQ_THREADS) ball_query_kernel(
    const float *__restrict__ new_xyz, const float *__restrict__ xyz, int N, int M, float radius2,
    int nsample, int32_t *__restrict__ idx_out) {
  extern __shared__ int32_t s_rows[];  // [BQ_WARPS*CPW][nsample]
  __shared__ float sx[BQ_TILE], sy[BQ_TILE], sz[BQ_TILE];
  __shared__ int s_active;

  const int scene = blockIdx.y;
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31u;
  const int warp = tid >> 5;
  const float *P = xyz + (size_t)scene * N * 3;
  const float *Qc = new_xyz + (size_t)scene * M * 3;
  const int c0 = (blockIdx.x * BQ_WARPS + warp) * CPW;  // first centre of this warp

  float cx[CPW], cy[CPW], cz[CPW];
  int cnt[CPW], first[CPW];
  int32_t *row[CPW];
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    const int j = c0 + c;
    const bool ok = j < M;
    cx[c] = ok ? __ldg(Qc + 3 * j + 0) : 0.f;
    cy[c] = ok ? __ldg(Qc + 3 * j + 1) : 0.f;
    cz[c] = ok ? __ldg(Qc + 3 * j + 2) : 0.f;
    cnt[c] = ok ? 0 : nsample;  // out-of-range centres count as "full"
    first[c] = 0;
    row[c] = s_rows + (size_t)(warp * CPW + c) * nsample;
  }
  if (tid == 0) s_active = 1;

  for (int base = 0; base < N; base += BQ_TILE) {
    __syncthreads();                       // previous tile fully consumed, s_active settled
    if (!s_active) break;                  // every warp of this CTA is done (uniform)
    __syncthreads();
    if (tid == 0) s_active = 0;
    const int tile = min(BQ_TILE, N - base);
    for (int e = tid; e < tile * 3; e += BQ_THREADS) {   // coalesced AoS read -> SoA smem
      const float v = __ldg(P + (size_t)base * 3 + e);
      const int pt = e / 3, comp = e - pt * 3;
      (comp == 0 ? sx : comp == 1 ? sy : sz)[pt] = v;
    }
    __syncthreads();
    bool warp_active = false;
#pragma unroll
    for (int c = 0; c < CPW; ++c) warp_active |= cnt[c] < nsample;
    if (warp_active) {
      for (int k0 = 0; k0 < tile; k0 += 32) {
        const int kk = k0 + lane;
        const bool inb = kk < tile;
        const float px = inb ? sx[kk] : 0.f, py = inb ? sy[kk] : 0.f, pz = inb ? sz[kk] : 0.f;
        bool any_open = false;
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
          const float d2 = sqdist_ref(cx[c], cy[c], cz[c], px, py, pz);
          const bool hit = inb && (d2 < radius2) && (cnt[c] < nsample);
          const unsigned mask = __ballot_sync(0xffffffffu, hit);
          if (mask) {
            if (cnt[c] == 0) first[c] = base + k0 + __ffs(mask) - 1;
            const int pos = cnt[c] + __popc(mask & ((1u << lane) - 1u));
            if (hit && pos < nsample) row[c][pos] = base + kk;
            cnt[c] = min(nsample, cnt[c] + __popc(mask));
          }
          any_open |= cnt[c] < nsample;
        }
        if (!any_open) { warp_active = false; break; }
      }
    }
    if (warp_active && lane == 0) s_active = 1;   // benign race: all writers store 1
  }
  __syncwarp();
  // ---- padding + coalesced write-back -----------------------------------------------------------
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    const int j = c0 + c;
    if (j >= M) continue;
    int32_t *dst = idx_out + ((size_t)scene * M + j) * nsample;
    for (int l = lane; l < nsample; l += 32) dst[l] = l < cnt[c] ? row[c][l] : first[c];
  }
}

}  // namespace spc

using namespace spc;

extern "C" int spc_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M,
                              float radius, int nsample, int32_t *idx, void *stream_) {
  SPC_CHECK_ARG(B >= 0 && N >= 0 && M >= 0 && nsample >= 0, "ball_query: bad sizes");
  if (B == 0 || M == 0 || nsample == 0) return SPC_OK;
  SPC_CHECK_ARG(new_xyz && idx && (xyz || N == 0), "ball_query: null pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  const float radius2 = radius * radius;  // f32 product, as ball_query_gpu.cu:22
  // centres per warp: 4 when there are plenty of centres, fewer to keep the grid wide otherwise
  int cpw = 4;
  while (cpw > 1 && (long long)B * ceil_div(M, BQ_WARPS * cpw) < 2 * kNumSMs) cpw >>= 1;
  const size_t smem = (size_t)BQ_WARPS * cpw * nsample * sizeof(int32_t);
  SPC_CHECK_ARG(smem <= 160 * 1024, "ball_query: nsample=%d too large", nsample);
  dim3 grid(ceil_div(M, BQ_WARPS * cpw), B);
#define BQ_LAUNCH(CPW)                                                                         \
  do {                                                                                         \
    if (smem > 20 * 1024)                                                                      \
      SPC_CUDA(cudaFuncSetAttribute(ball_query_kernel<CPW>,                                    \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    ball_query_kernel<CPW><<<grid, BQ_THREADS, smem, stream>>>(new_xyz, xyz, N, M, radius2,    \
                                                               nsample, idx);                  \
  } while (0)
  if (cpw == 4) BQ_LAUNCH(4);
  else if (cpw == 2) BQ_LAUNCH(2);
  else BQ_LAUNCH(1);
  SPC_LAUNCH_CHECK("ball_query_kernel");
  return SPC_OK;
}
