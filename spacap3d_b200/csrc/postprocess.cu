// postprocess.cu -- device side of parse_predictions (reference lib/ap_helper.py:44-160):
//   * spc_box_point_counts : "remove predicted boxes with fewer than 5 points" (ap_helper.py:69-79), which the
//                            reference does with B*K scipy Delaunay hull tests over the whole cloud on the host;
//   * spc_nms_boxes        : nms_2d_faster / nms_3d_faster / nms_3d_faster_samecls (utils/nms.py:39-147), which the
//                            reference runs per scene in numpy on the host.
// Arithmetic is fp64 like numpy's (the box arrays there are float64), same expression order, so the picks are
// identical whenever the scores are distinct; equal scores are ordered as by a STABLE ascending sort (numpy's
// default introsort gives no guarantee there).
#include "common.cuh"

namespace spc {

constexpr int PP_THREADS = 256;
constexpr int PP_BOXES = 32;        // boxes per CTA in the point-count kernel
constexpr int NMS_MAX_K = 512;     // 36 KB of static shared memory

// corners (8,3) fp64 in the order of get_3d_box_batch (utils/box_util.py:360-383): edges from corner 0 run to
// corners 1, 3 and 4.  A point is inside the (possibly rotated) box iff its three edge coordinates are in [0,1].
struct BoxFrame {
  double o[3], e[3][3], inv[3];     // origin, edge vectors, 1 / |edge|^2
};

__global__ void __launch_bounds__(PP_THREADS) box_point_count_kernel(const float *__restrict__ pts, int pt_stride, int N,
                                                                      const double *__restrict__ corners, int K,
                                                                      int32_t *__restrict__ counts) {
  __shared__ BoxFrame s_box[PP_BOXES];
  __shared__ int s_cnt[PP_BOXES];
  const int b = blockIdx.y;
  const int k0 = blockIdx.x * PP_BOXES;
  const int nb = min(PP_BOXES, K - k0);
  if (threadIdx.x < nb) {
    const double *c = corners + ((size_t)b * K + k0 + threadIdx.x) * 24;
    BoxFrame f;
    const int other[3] = {1, 3, 4};
    for (int d = 0; d < 3; ++d) f.o[d] = c[d];
    for (int a = 0; a < 3; ++a) {
      double n2 = 0.0;
      for (int d = 0; d < 3; ++d) { f.e[a][d] = c[other[a] * 3 + d] - c[d]; n2 += f.e[a][d] * f.e[a][d]; }
      f.inv[a] = n2 > 0.0 ? 1.0 / n2 : 0.0;
    }
    s_box[threadIdx.x] = f;
    s_cnt[threadIdx.x] = 0;
  }
  __syncthreads();
  int local[PP_BOXES];
#pragma unroll
  for (int q = 0; q < PP_BOXES; ++q) local[q] = 0;
  const float *P = pts + (size_t)b * N * pt_stride;
  for (int i = blockIdx.z * PP_THREADS + threadIdx.x; i < N; i += gridDim.z * PP_THREADS) {
    const double x = (double)__ldg(P + (size_t)i * pt_stride), y = (double)__ldg(P + (size_t)i * pt_stride + 1),
                 z = (double)__ldg(P + (size_t)i * pt_stride + 2);
#pragma unroll
    for (int q = 0; q < PP_BOXES; ++q) {
      if (q < nb) {
        const BoxFrame &f = s_box[q];
        const double dx = x - f.o[0], dy = y - f.o[1], dz = z - f.o[2];
        bool in = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const double t = (dx * f.e[a][0] + dy * f.e[a][1] + dz * f.e[a][2]) * f.inv[a];
          in = in && t >= 0.0 && t <= 1.0;
        }
        local[q] += in ? 1 : 0;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < PP_BOXES; ++q) {
    int v = local[q];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[q], v);
  }
  __syncthreads();
  if (threadIdx.x < nb && s_cnt[threadIdx.x]) atomicAdd(counts + (size_t)b * K + k0 + threadIdx.x, s_cnt[threadIdx.x]);
}

// one CTA per scene.  mode: 0 = 2-D (x and z extents, utils/nms.py:39-70), 1 = 3-D (:72-107), 2 = 3-D suppressing
// only boxes of the same class (:109-147, the variant SpaCap3D evaluates with, scripts/eval.py:195-203).
__global__ void __launch_bounds__(PP_THREADS) nms_boxes_kernel(const double *__restrict__ corners,
                                                                const float *__restrict__ score,
                                                                const int64_t *__restrict__ cls,
                                                                const int32_t *__restrict__ valid, int K, int mode,
                                                                int old_type, double thr, int32_t *__restrict__ pick) {
  __shared__ double s_lo[NMS_MAX_K][3], s_hi[NMS_MAX_K][3], s_area[NMS_MAX_K], s_score[NMS_MAX_K];
  __shared__ int s_order[NMS_MAX_K], s_cls[NMS_MAX_K];
  __shared__ unsigned char s_state[NMS_MAX_K];       // 0 = candidate, 1 = suppressed / invalid, 2 = picked
  __shared__ int s_nvalid, s_cur;
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) s_nvalid = 0;
  __syncthreads();
  for (int j = tid; j < K; j += PP_THREADS) {
    const double *c = corners + ((size_t)b * K + j) * 24;
    double lo[3] = {c[0], c[1], c[2]}, hi[3] = {c[0], c[1], c[2]};
    for (int q = 1; q < 8; ++q)
      for (int d = 0; d < 3; ++d) { lo[d] = fmin(lo[d], c[q * 3 + d]); hi[d] = fmax(hi[d], c[q * 3 + d]); }
    for (int d = 0; d < 3; ++d) { s_lo[j][d] = lo[d]; s_hi[j][d] = hi[d]; }
    // numpy: (x2-x1)*(y2-y1)*(z2-z1) left to right; the 2-D variant uses the x and z extents
    s_area[j] = mode == 0 ? (hi[0] - lo[0]) * (hi[2] - lo[2]) : ((hi[0] - lo[0]) * (hi[1] - lo[1])) * (hi[2] - lo[2]);
    s_score[j] = (double)score[(size_t)b * K + j];
    s_cls[j] = cls ? (int)cls[(size_t)b * K + j] : 0;
    const bool v = valid == nullptr || valid[(size_t)b * K + j] != 0;
    s_state[j] = v ? 0 : 1;
    if (v) atomicAdd(&s_nvalid, 1);
  }
  __syncthreads();
  // rank among valid boxes by (score, index) ascending = position in a stable argsort
  for (int j = tid; j < K; j += PP_THREADS) {
    if (s_state[j] != 0) continue;
    const double sj = s_score[j];
    int r = 0;
    for (int i = 0; i < K; ++i) r += (s_state[i] == 0 && (s_score[i] < sj || (s_score[i] == sj && i < j))) ? 1 : 0;
    s_order[r] = j;
  }
  __syncthreads();
  const int nvalid = s_nvalid;
  for (int r = nvalid - 1; r >= 0; --r) {
    if (tid == 0) {
      const int i = s_order[r];
      if (s_state[i] == 0) { s_state[i] = 2; s_cur = i; } else s_cur = -1;
    }
    __syncthreads();
    const int i = s_cur;
    if (i >= 0) {
      for (int j = tid; j < K; j += PP_THREADS) {
        if (s_state[j] != 0) continue;
        double o;
        if (mode == 0) {
          const double w = fmax(0.0, fmin(s_hi[i][0], s_hi[j][0]) - fmax(s_lo[i][0], s_lo[j][0]));
          const double h = fmax(0.0, fmin(s_hi[i][2], s_hi[j][2]) - fmax(s_lo[i][2], s_lo[j][2]));
          const double inter = w * h;
          o = old_type ? inter / s_area[j] : inter / (s_area[i] + s_area[j] - inter);
        } else {
          const double l = fmax(0.0, fmin(s_hi[i][0], s_hi[j][0]) - fmax(s_lo[i][0], s_lo[j][0]));
          const double w = fmax(0.0, fmin(s_hi[i][1], s_hi[j][1]) - fmax(s_lo[i][1], s_lo[j][1]));
          const double h = fmax(0.0, fmin(s_hi[i][2], s_hi[j][2]) - fmax(s_lo[i][2], s_lo[j][2]));
          const double inter = (l * w) * h;
          if (old_type) o = inter / s_area[j];
          else if (mode == 2) o = inter / (((s_area[i] + s_area[j]) - inter) + 1e-8);
          else o = inter / ((s_area[i] + s_area[j]) - inter);
          if (mode == 2) o = o * (s_cls[i] == s_cls[j] ? 1.0 : 0.0);
        }
        if (o > thr) s_state[j] = 1;
      }
    }
    __syncthreads();
  }
  for (int j = tid; j < K; j += PP_THREADS) pick[(size_t)b * K + j] = s_state[j] == 2 ? 1 : 0;
}

}  // namespace spc

using namespace spc;

extern "C" int spc_box_point_counts(const float *points, int point_stride, const double *corners, int B, int N, int K,
                                    int32_t *counts, void *stream_) {
  SPC_CHECK_ARG(B >= 0 && N >= 0 && K >= 0 && point_stride >= 3, "box_point_counts: bad sizes");
  if (B == 0 || K == 0) return SPC_OK;
  SPC_CHECK_ARG(points && corners && counts, "box_point_counts: null pointer");
  SPC_CHECK_ARG(B <= 65535, "box_point_counts: B too large");
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_CUDA(cudaMemsetAsync(counts, 0, (size_t)B * K * sizeof(int32_t), stream));
  if (N == 0) return SPC_OK;
  const int tiles = ceil_div(K, PP_BOXES);
  int zsplit = max(1, min(64, ceil_div(2 * kNumSMs, tiles * B)));
  zsplit = min(zsplit, ceil_div(N, PP_THREADS));
  box_point_count_kernel<<<dim3(tiles, B, zsplit), PP_THREADS, 0, stream>>>(points, point_stride, N, corners, K, counts);
  SPC_LAUNCH_CHECK("box_point_count_kernel");
  return SPC_OK;
}

extern "C" int spc_nms_boxes(const double *corners, const float *score, const int64_t *cls, const int32_t *valid, int B,
                             int K, int mode, int old_type, double iou_threshold, int32_t *pick, void *stream_) {
  SPC_CHECK_ARG(B >= 0 && K >= 0 && mode >= 0 && mode <= 2, "nms_boxes: bad arguments");
  if (B == 0 || K == 0) return SPC_OK;
  SPC_CHECK_ARG(corners && score && pick && (mode != 2 || cls), "nms_boxes: null pointer");
  if (K > NMS_MAX_K) {
    set_error("nms_boxes: K=%d exceeds %d proposals per scene", K, NMS_MAX_K);
    return SPC_ERR_UNSUPPORTED;
  }
  nms_boxes_kernel<<<B, PP_THREADS, 0, (cudaStream_t)stream_>>>(corners, score, cls, valid, K, mode, old_type,
                                                                iou_threshold, pick);
  SPC_LAUNCH_CHECK("nms_boxes_kernel");
  return SPC_OK;
}
