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
#include <stdlib.h>

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


// ================================================================================================
// Grid-accelerated ball query (large clouds).
//
// The brute-force kernel above tests every (centre, point) pair: 655 M tests for SA1 at batch 8,
// of which only ~37 per centre hit.  Here the scene's points are binned into a uniform grid whose
// cell edge is >= radius, so a centre only needs the 3x3x3 cells around it (3 contiguous runs of
// cells per (y,z) row => 9 ranges).  The reference's contract -- the FIRST nsample hits in
// ascending point index, padded with the first -- is kept exactly: the hit test is the same
// sqdist_ref(...) < r*r on the same operands, the candidate set provably contains every hit
// (cell edge = r*(1+1e-4) absorbs the rounding of the cell-index computation, which is monotone),
// and the hits are then ordered by index with a rank-by-counting pass.  Centres with more than
// BQG_CAP hits fall back to the ordered brute-force scan.
// ================================================================================================
constexpr int BQG_MAX_CELLS = 32768;
constexpr int BQG_BUILD_THREADS = 1024;
constexpr int BQG_CAP = 512;          // hits buffered per centre
constexpr int BQG_THREADS = 256;      // query kernel: 8 warps = 8 centres per CTA

struct BqGrid {        // per scene, written by the build kernel
  float minx, miny, minz, inv_h;
  int gx, gy, gz, ncell;
};

__device__ __forceinline__ int bqg_axis_cell(float v, float mn, float inv_h, int g) {
  // monotone in v; clamped so that far-away centres map to "one past the border"
  const float f = floorf(__fmul_rn(__fsub_rn(v, mn), inv_h));
  return (int)fminf(fmaxf(f, -2.0f), (float)(g + 1));
}

// one CTA per scene: bbox -> grid parameters -> counting sort of the points by cell
__global__ void __launch_bounds__(BQG_BUILD_THREADS) bqg_build_kernel(const float *__restrict__ xyz, int N,
                                                                      float radius, BqGrid *__restrict__ grids,
                                                                      int *__restrict__ cell_start,
                                                                      float4 *__restrict__ sorted_pts) {
  extern __shared__ int s_cnt[];                 // [BQG_MAX_CELLS]
  __shared__ float s_red[6][32];
  __shared__ BqGrid s_g;
  __shared__ int s_part[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *P = xyz + (size_t)b * N * 3;
  // ---- bounding box -------------------------------------------------------------------------
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int k = tid; k < N; k += BQG_BUILD_THREADS) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __ldg(P + 3 * k + c);
      mn[c] = fminf(mn[c], v);
      mx[c] = fmaxf(mx[c], v);
    }
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
  if (tid == 0) {
    float lo[3], hi[3];
    for (int c = 0; c < 3; ++c) {
      lo[c] = s_red[c][0]; hi[c] = s_red[3 + c][0];
      for (int w = 1; w < BQG_BUILD_THREADS / 32; ++w) { lo[c] = fminf(lo[c], s_red[c][w]); hi[c] = fmaxf(hi[c], s_red[3 + c][w]); }
    }
    float h = radius * 1.0001f;                  // cell edge >= radius with a rounding margin
    if (!(h > 0.f)) h = 1.0f;
    int g[3];
    for (;;) {
      long long tot = 1;
      for (int c = 0; c < 3; ++c) {
        const float e = fmaxf(hi[c] - lo[c], 0.f);
        const float cells = floorf(e / h) + 1.0f;
        g[c] = cells > 1e6f ? 1000000 : (int)cells;
        tot *= g[c];
      }
      if (tot <= BQG_MAX_CELLS) break;
      h *= 1.26f;                                // coarser cells stay correct, just less selective
    }
    s_g.minx = lo[0]; s_g.miny = lo[1]; s_g.minz = lo[2]; s_g.inv_h = 1.0f / h;
    s_g.gx = g[0]; s_g.gy = g[1]; s_g.gz = g[2]; s_g.ncell = g[0] * g[1] * g[2];
    grids[b] = s_g;
  }
  __syncthreads();
  const BqGrid G = s_g;
  for (int c = tid; c < G.ncell; c += BQG_BUILD_THREADS) s_cnt[c] = 0;
  __syncthreads();
  // ---- histogram ------------------------------------------------------------------------------
  for (int k = tid; k < N; k += BQG_BUILD_THREADS) {
    const int cx = min(max(bqg_axis_cell(__ldg(P + 3 * k), G.minx, G.inv_h, G.gx), 0), G.gx - 1);
    const int cy = min(max(bqg_axis_cell(__ldg(P + 3 * k + 1), G.miny, G.inv_h, G.gy), 0), G.gy - 1);
    const int cz = min(max(bqg_axis_cell(__ldg(P + 3 * k + 2), G.minz, G.inv_h, G.gz), 0), G.gz - 1);
    atomicAdd(&s_cnt[(cz * G.gy + cy) * G.gx + cx], 1);
  }
  __syncthreads();
  // ---- exclusive scan over cells (each thread owns a contiguous run) --------------------------
  const int per = (G.ncell + BQG_BUILD_THREADS - 1) / BQG_BUILD_THREADS;
  const int c0 = tid * per, c1 = min(G.ncell, c0 + per);
  int sum = 0;
  for (int c = c0; c < c1; ++c) sum += s_cnt[c];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_part[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = s_part[lane], inc2 = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc2, o); if (lane >= o) inc2 += u; }
    s_part[lane] = inc2 - v;                     // exclusive prefix of the warp totals
  }
  __syncthreads();
  int run = s_part[warp] + incl - sum;           // exclusive prefix of this thread's run
  int *cs = cell_start + (size_t)b * (BQG_MAX_CELLS + 1);
  for (int c = c0; c < c1; ++c) { const int n = s_cnt[c]; s_cnt[c] = run; cs[c] = run; run += n; }
  if (tid == 0) cs[G.ncell] = N;
  __syncthreads();
  // ---- scatter (order inside a cell is arbitrary; hits are ordered by index later) -----------
  float4 *out = sorted_pts + (size_t)b * N;
  for (int k = tid; k < N; k += BQG_BUILD_THREADS) {
    const float x = __ldg(P + 3 * k), y = __ldg(P + 3 * k + 1), z = __ldg(P + 3 * k + 2);
    const int cx = min(max(bqg_axis_cell(x, G.minx, G.inv_h, G.gx), 0), G.gx - 1);
    const int cy = min(max(bqg_axis_cell(y, G.miny, G.inv_h, G.gy), 0), G.gy - 1);
    const int cz = min(max(bqg_axis_cell(z, G.minz, G.inv_h, G.gz), 0), G.gz - 1);
    const int pos = atomicAdd(&s_cnt[(cz * G.gy + cy) * G.gx + cx], 1);
    out[pos] = make_float4(x, y, z, __int_as_float(k));
  }
}

// one warp per centre
__global__ void __launch_bounds__(BQG_THREADS) bqg_query_kernel(const float *__restrict__ new_xyz,
                                                                const float *__restrict__ xyz, int N, int M,
                                                                float radius, float radius2, int nsample,
                                                                const BqGrid *__restrict__ grids,
                                                                const int *__restrict__ cell_start,
                                                                const float4 *__restrict__ sorted_pts,
                                                                int32_t *__restrict__ idx_out) {
  extern __shared__ int32_t s_buf[];             // [8 warps][BQG_CAP hits] + [8 warps][nsample] rows
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5;
  const unsigned lane = tid & 31u;
  const int j = blockIdx.x * (BQG_THREADS / 32) + warp;
  if (j >= M) return;                            // whole warp leaves together
  int32_t *hits = s_buf + warp * BQG_CAP;
  int32_t *row = s_buf + (BQG_THREADS / 32) * BQG_CAP + warp * nsample;
  const BqGrid G = grids[b];
  const int *cs = cell_start + (size_t)b * (BQG_MAX_CELLS + 1);
  const float4 *pts = sorted_pts + (size_t)b * N;
  const float *Q = new_xyz + ((size_t)b * M + j) * 3;
  const float qx = __ldg(Q), qy = __ldg(Q + 1), qz = __ldg(Q + 2);
  // Candidate cells: [cell(q - r'), cell(q + r')] per axis with r' a hair above the radius.  bqg_axis_cell is a
  // monotone function of the coordinate (fp32 subtraction, multiplication by a positive number and floor all are),
  // and every hit has |p.x - q.x| <= |p - q| < r (1 + ~1e-7), so its cell lies inside that range WHATEVER the
  // rounding of the cell arithmetic -- also for clouds spanning hundreds of cells, where "own cell +- 1" could miss a
  // neighbour two cells away (round-1 advisor finding).  The slack covers the rounding of q -+ r' itself.
  const float rx = fmaf(fabsf(qx), 4e-7f, radius * 1.0001f), ry = fmaf(fabsf(qy), 4e-7f, radius * 1.0001f),
              rz = fmaf(fabsf(qz), 4e-7f, radius * 1.0001f);
  const int x0 = max(bqg_axis_cell(qx - rx, G.minx, G.inv_h, G.gx), 0),
            x1 = min(bqg_axis_cell(qx + rx, G.minx, G.inv_h, G.gx), G.gx - 1);
  const int y0 = max(bqg_axis_cell(qy - ry, G.miny, G.inv_h, G.gy), 0),
            y1 = min(bqg_axis_cell(qy + ry, G.miny, G.inv_h, G.gy), G.gy - 1);
  const int z0 = max(bqg_axis_cell(qz - rz, G.minz, G.inv_h, G.gz), 0),
            z1 = min(bqg_axis_cell(qz + rz, G.minz, G.inv_h, G.gz), G.gz - 1);
  int cnt = 0;
  bool overflow = false;
  if (x0 <= x1) {
    for (int zz = z0; zz <= z1; ++zz) {
      for (int yy = y0; yy <= y1; ++yy) {
        const int rowbase = (zz * G.gy + yy) * G.gx;
        const int beg = __ldg(cs + rowbase + x0), end = __ldg(cs + rowbase + x1 + 1);
        for (int t0 = beg; t0 < end; t0 += 32) {
          const int t = t0 + (int)lane;
          bool hit = false;
          int k = 0;
          if (t < end) {
            const float4 p = __ldg(pts + t);
            k = __float_as_int(p.w);
            hit = sqdist_ref(qx, qy, qz, p.x, p.y, p.z) < radius2;
          }
          const unsigned mask = __ballot_sync(0xffffffffu, hit);
          if (mask) {
            const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
            if (hit && pos < BQG_CAP) hits[pos] = k;
            cnt += __popc(mask);
          }
        }
      }
    }
  }
  overflow = cnt > BQG_CAP;
  __syncwarp();
  int32_t *dst = idx_out + ((size_t)b * M + j) * nsample;
  if (!overflow) {
    // ---- order the hits by point index: rank = number of hits with a smaller index -----------
    unsigned first = 0x7fffffffu;
    for (int i = lane; i < cnt; i += 32) first = min(first, (unsigned)hits[i]);
    first = __reduce_min_sync(0xffffffffu, first);
    if (cnt == 0) first = 0u;                    // empty ball: all-zero row (F9)
    for (int l = lane; l < nsample; l += 32) row[l] = (int)first;   // padding
    __syncwarp();
    for (int i = lane; i < cnt; i += 32) {
      const int mine = hits[i];
      int rank = 0;
      for (int o = 0; o < cnt; ++o) rank += hits[o] < mine;          // broadcast smem reads
      if (rank < nsample) row[rank] = mine;
    }
    __syncwarp();
    for (int l = lane; l < nsample; l += 32) dst[l] = row[l];
    return;
  }
  // ---- fallback: ordered brute-force scan of the original array for this centre ---------------
  const float *P = xyz + (size_t)b * N * 3;
  int c2 = 0, first = 0;
  for (int k0 = 0; k0 < N && c2 < nsample; k0 += 32) {
    const int kk = k0 + (int)lane;
    bool hit = false;
    if (kk < N) hit = sqdist_ref(qx, qy, qz, __ldg(P + 3 * kk), __ldg(P + 3 * kk + 1), __ldg(P + 3 * kk + 2)) < radius2;
    const unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (mask) {
      if (c2 == 0) first = k0 + __ffs(mask) - 1;
      const int pos = c2 + __popc(mask & ((1u << lane) - 1u));
      if (hit && pos < nsample) row[pos] = kk;
      c2 = min(nsample, c2 + __popc(mask));
    }
  }
  __syncwarp();
  for (int l = lane; l < nsample; l += 32) dst[l] = l < c2 ? row[l] : first;
}

}  // namespace spc

using namespace spc;

extern "C" size_t spc_ball_query_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return (size_t)B * (sizeof(BqGrid) + (size_t)(BQG_MAX_CELLS + 1) * 4 + (size_t)N * sizeof(float4)) + 256;
}

static int ball_query_brute(const float *new_xyz, const float *xyz, int B, int N, int M, float radius,
                            int nsample, int32_t *idx, void *stream_);

extern "C" int spc_ball_query_ex(const float *new_xyz, const float *xyz, int B, int N, int M,
                                 float radius, int nsample, int32_t *idx, void *workspace,
                                 size_t workspace_bytes, void *stream_) {
  SPC_CHECK_ARG(B >= 0 && N >= 0 && M >= 0 && nsample >= 0, "ball_query: bad sizes");
  if (B == 0 || M == 0 || nsample == 0) return SPC_OK;
  SPC_CHECK_ARG(new_xyz && idx && (xyz || N == 0), "ball_query: null pointer");
  // the grid pays off once the all-pairs scan dominates; tiny clouds stay on the brute-force kernel
  const bool use_grid = workspace && workspace_bytes >= spc_ball_query_workspace_bytes(B, N) &&
                        N >= 1024 && (long long)N * M >= (1LL << 19) && nsample <= 1024 && B <= 65535 &&
                        radius > 0.f;
  if (!use_grid) return ball_query_brute(new_xyz, xyz, B, N, M, radius, nsample, idx, stream_);
  cudaStream_t stream = (cudaStream_t)stream_;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  float4 *sorted_pts = reinterpret_cast<float4 *>(base);
  int *cell_start = reinterpret_cast<int *>(sorted_pts + (size_t)B * N);
  BqGrid *grids = reinterpret_cast<BqGrid *>(cell_start + (size_t)B * (BQG_MAX_CELLS + 1));
  const size_t build_smem = (size_t)BQG_MAX_CELLS * sizeof(int);
  SPC_CUDA(cudaFuncSetAttribute(bqg_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_smem));
  bqg_build_kernel<<<B, BQG_BUILD_THREADS, build_smem, stream>>>(xyz, N, radius, grids, cell_start, sorted_pts);
  SPC_LAUNCH_CHECK("bqg_build_kernel");
  const size_t q_smem = (size_t)(BQG_THREADS / 32) * (BQG_CAP + nsample) * sizeof(int32_t);
  if (q_smem > 40 * 1024)
    SPC_CUDA(cudaFuncSetAttribute(bqg_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q_smem));
  const float radius2 = radius * radius;   // f32 product, as ball_query_gpu.cu:22
  bqg_query_kernel<<<dim3(ceil_div(M, BQG_THREADS / 32), B), BQG_THREADS, q_smem, stream>>>(
      new_xyz, xyz, N, M, radius, radius2, nsample, grids, cell_start, sorted_pts, idx);
  SPC_LAUNCH_CHECK("bqg_query_kernel");
  return SPC_OK;
}

extern "C" int spc_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M,
                              float radius, int nsample, int32_t *idx, void *stream_) {
  return ball_query_brute(new_xyz, xyz, B, N, M, radius, nsample, idx, stream_);
}

static int ball_query_brute(const float *new_xyz, const float *xyz, int B, int N, int M, float radius,
                            int nsample, int32_t *idx, void *stream_) {
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
