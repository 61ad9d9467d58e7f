/*
 * pointnet2_oracle.c -- CPU restatement of the reference's nine PointNet++ CUDA kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker*: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may call it.  The product path
 * (spacap3d_b200/) never links, imports or falls back to it.
 *
 * Parity status: PINNED.  tests/golden/ref_ops_*.npz hold outputs of the reference's own CUDA
 * extension (rebuilt unmodified for sm_100a by oracle/build_ref.py, run on a B200 by
 * oracle/make_golden.py); tests/test_oracle_golden.py checks every function below against them
 * bit-for-bit (indices, copies) or to 1e-6 (atomically accumulated gradients).
 *
 * Every function restates one reference kernel; citations are into
 * /root/reference/lib/pointnet2/_ext_src/.  Rounding contract: nvcc contracts
 *   (a*a) + (b*b) + (c*c)    into   fmaf(c, c, fmaf(a, a, b*b))
 * so the squared distances below are written with explicit fmaf in that order and the file is
 * compiled with -ffp-contract=off so that gcc adds no contraction of its own.
 *
 * Build: see oracle/Makefile (gcc -O2 -mfma -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_TOTAL_THREADS 512

/* include/cuda_utils.h:15-19 -- largest power of two <= work_size, capped at 512. */
int orc_opt_n_threads(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int t = 1 << pow_2;
  if (t > ORC_TOTAL_THREADS) t = ORC_TOTAL_THREADS;
  if (t < 1) t = 1;
  return t;
}

/* include/cuda_utils.h:21-28 */
void orc_opt_block_config(int x, int y, int *bx, int *by) {
  const int xt = orc_opt_n_threads(x);
  int yt = orc_opt_n_threads(y);
  if (yt > ORC_TOTAL_THREADS / xt) yt = ORC_TOTAL_THREADS / xt;
  if (yt < 1) yt = 1;
  *bx = xt;
  *by = yt;
}

/* squared distance exactly as the SASS of sampling_gpu.cu:103-104, ball_query_gpu.cu:31-32 and
 * interpolate_gpu.cu:33 evaluates it: FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.)            */
static inline float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* ------------------------------------------------------------------------------------------
 * furthest_point_sampling_kernel, sampling_gpu.cu:69-173 (+ __update :59-65, launch :175-229,
 * temp = 1e10 and idx = zeros from sampling.cpp:70-76).
 * dataset (b,n,3) -> idxs (b,m).  The block of T = opt_n_threads(n) threads is simulated
 * faithfully: per-thread strided scan with strict '>' and the shared-memory tree whose
 * "v2 > v1 ? i2 : i1" keeps the LEFT operand on ties.
 * ------------------------------------------------------------------------------------------ */
void orc_furthest_point_sampling(int b, int n, int m, const float *dataset, int *idxs) {
  if (m <= 0) return;
  const int T = orc_opt_n_threads(n);
#pragma omp parallel for schedule(dynamic, 1)
  for (int bi = 0; bi < b; ++bi) {
    const float *p = dataset + (size_t)bi * n * 3;
    int *out = idxs + (size_t)bi * m;
    float *temp = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    float *dists = (float *)malloc(sizeof(float) * T);
    int *dists_i = (int *)malloc(sizeof(int) * T);
    for (int k = 0; k < n; ++k) temp[k] = 1e10f;
    memset(out, 0, sizeof(int) * (size_t)m);
    int old = 0;
    out[0] = old;
    for (int j = 1; j < m; ++j) {
      const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
      for (int tid = 0; tid < T; ++tid) {
        int besti = 0;
        float best = -1.0f;
        for (int k = tid; k < n; k += T) {
          const float x2 = p[k * 3 + 0], y2 = p[k * 3 + 1], z2 = p[k * 3 + 2];
          /* :100 mag = x2*x2 + y2*y2 + z2*z2, contracted as fmaf(z,z,fmaf(x,x,y*y)) */
          const float mag = fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
          if ((double)mag <= 1e-3) continue; /* :101, compare done in double (F5) */
          const float d = sqdist(x2, y2, z2, x1, y1, z1);
          const float d2 = fminf(d, temp[k]);
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        dists[tid] = best;
        dists_i[tid] = besti;
      }
      for (int s = T / 2; s >= 1; s >>= 1) { /* :115-168 */
        for (int tid = 0; tid < s; ++tid) {
          const float v1 = dists[tid], v2 = dists[tid + s];
          const int i1 = dists_i[tid], i2 = dists_i[tid + s];
          dists[tid] = v1 > v2 ? v1 : (v2 > v1 ? v2 : v1); /* max(v1,v2) */
          dists_i[tid] = v2 > v1 ? i2 : i1;
        }
      }
      old = dists_i[0];
      out[j] = old;
    }
    free(temp);
    free(dists);
    free(dists_i);
  }
}

/* gather_points_kernel, sampling_gpu.cu:8-20.  points (b,c,n), idx (b,m) -> out (b,c,m) */
void orc_gather_points(int b, int c, int n, int m, const float *points, const int *idx,
                       float *out) {
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j) {
        const int a = idx[(size_t)i * m + j];
        out[((size_t)i * c + l) * m + j] = points[((size_t)i * c + l) * n + a];
      }
}

/* gather_points_grad_kernel, sampling_gpu.cu:34-47.  grad_points (b,c,n) zeroed first
 * (sampling.cpp:51-53), then scatter-add (sequential order here; the GPU order is arbitrary). */
void orc_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                            float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j) {
        const int a = idx[(size_t)i * m + j];
        grad_points[((size_t)i * c + l) * n + a] += grad_out[((size_t)i * c + l) * m + j];
      }
}

/* query_ball_point_kernel, ball_query_gpu.cu:9-44; idx zero-filled by ball_query.cpp:19-21.
 * new_xyz (b,m,3), xyz (b,n,3) -> idx (b,m,nsample) */
void orc_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                    const float *xyz, int *idx) {
  const float radius2 = radius * radius; /* :22, f32 product */
  memset(idx, 0, sizeof(int) * (size_t)b * m * nsample);
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi) {
    for (int j = 0; j < m; ++j) {
      const float *P = xyz + (size_t)bi * n * 3;
      const float *Q = new_xyz + (size_t)bi * m * 3;
      int *o = idx + ((size_t)bi * m + j) * nsample;
      const float nx = Q[j * 3 + 0], ny = Q[j * 3 + 1], nz = Q[j * 3 + 2];
      for (int k = 0, cnt = 0; k < n && cnt < nsample; ++k) {
        const float d2 = sqdist(nx, ny, nz, P[k * 3 + 0], P[k * 3 + 1], P[k * 3 + 2]);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) o[l] = k;
          o[cnt] = k;
          ++cnt;
        }
      }
    }
  }
}

/* group_points_kernel, group_points_gpu.cu:8-28.  points (b,c,n), idx (b,np,ns) -> (b,c,np,ns) */
void orc_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                      const int *idx, float *out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      const float *src = points + ((size_t)bi * c + l) * n;
      const int *ix = idx + (size_t)bi * npoints * nsample;
      float *dst = out + ((size_t)bi * c + l) * npoints * nsample;
      for (size_t t = 0; t < (size_t)npoints * nsample; ++t) dst[t] = src[ix[t]];
    }
}

/* group_points_grad_kernel, group_points_gpu.cu:43-64 (zero fill group_points.cpp:48-50). */
void orc_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                           const int *idx, float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      float *dst = grad_points + ((size_t)bi * c + l) * n;
      const int *ix = idx + (size_t)bi * npoints * nsample;
      const float *g = grad_out + ((size_t)bi * c + l) * npoints * nsample;
      for (size_t t = 0; t < (size_t)npoints * nsample; ++t) dst[ix[t]] += g[t];
    }
}

/* three_nn_kernel, interpolate_gpu.cu:9-59.  unknown (b,n,3), known (b,m,3) ->
 * dist2 (b,n,3) f32, idx (b,n,3).  Running bests are double initialised to 1e40 (:27). */
void orc_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                  int *idx) {
  for (int bi = 0; bi < b; ++bi) {
    const float *U = unknown + (size_t)bi * n * 3;
    const float *K = known + (size_t)bi * m * 3;
    for (int j = 0; j < n; ++j) {
      const float ux = U[j * 3 + 0], uy = U[j * 3 + 1], uz = U[j * 3 + 2];
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int besti1 = 0, besti2 = 0, besti3 = 0;
      for (int k = 0; k < m; ++k) {
        const float d = sqdist(ux, uy, uz, K[k * 3 + 0], K[k * 3 + 1], K[k * 3 + 2]);
        if (d < best1) {
          best3 = best2; besti3 = besti2;
          best2 = best1; besti2 = besti1;
          best1 = d; besti1 = k;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2;
          best2 = d; besti2 = k;
        } else if (d < best3) {
          best3 = d; besti3 = k;
        }
      }
      float *od = dist2 + ((size_t)bi * n + j) * 3;
      int *oi = idx + ((size_t)bi * n + j) * 3;
      od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
      oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
    }
  }
}

/* three_interpolate_kernel, interpolate_gpu.cu:72-101.  points (b,c,m), idx/weight (b,n,3)
 * -> out (b,c,n).  "p1*w1 + p2*w2 + p3*w3" is contracted by nvcc to
 * fmaf(p3,w3, fmaf(p1,w1, p2*w2)) (same pattern as F3). */
void orc_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                           const float *weight, float *out) {
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      const float *src = points + ((size_t)bi * c + l) * m;
      float *dst = out + ((size_t)bi * c + l) * n;
      for (int j = 0; j < n; ++j) {
        const float *w = weight + ((size_t)bi * n + j) * 3;
        const int *ix = idx + ((size_t)bi * n + j) * 3;
        dst[j] = fmaf(src[ix[2]], w[2], fmaf(src[ix[0]], w[0], src[ix[1]] * w[1]));
      }
    }
}

/* three_interpolate_grad_kernel, interpolate_gpu.cu:116-143 (zero fill interpolate.cpp:83-85) */
void orc_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                const float *weight, float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * c * m);
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      const float *g = grad_out + ((size_t)bi * c + l) * n;
      float *dst = grad_points + ((size_t)bi * c + l) * m;
      for (int j = 0; j < n; ++j) {
        const float *w = weight + ((size_t)bi * n + j) * 3;
        const int *ix = idx + ((size_t)bi * n + j) * 3;
        dst[ix[0]] += g[j] * w[0];
        dst[ix[1]] += g[j] * w[1];
        dst[ix[2]] += g[j] * w[2];
      }
    }
}
