// Latency microbenchmarks (one warp, dependent chains) for the primitives used in the FPS
// reduction: redux.sync, shuffle butterfly, ballot, shared-memory round trips.  Developer tool.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench_lat tools/microbench_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__global__ void k(unsigned *out, long long *cyc, int nwarps_active) {
  __shared__ unsigned sm[1024];
  unsigned v = threadIdx.x * 2654435761u + 12345u;
  const unsigned lane = threadIdx.x & 31;
  long long t0, t1;
  int slot = 0;
  // 1. redux.max chain
  t0 = clock64();
  for (int i = 0; i < N; ++i) { v = __reduce_max_sync(0xffffffffu, v ^ lane) + i; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / N; slot++;
  // 2. butterfly max with shfl_xor (5 steps)
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    unsigned w = v ^ lane;
#pragma unroll
    for (int o = 16; o; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    v = w + i;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / N; slot++;
  // 3. ballot + ffs chain
  t0 = clock64();
  for (int i = 0; i < N; ++i) { unsigned b = __ballot_sync(0xffffffffu, (v + lane) & 1); v = v + __ffs(b) + i; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / N; slot++;
  // 4. shfl idx chain
  t0 = clock64();
  for (int i = 0; i < N; ++i) { v = __shfl_sync(0xffffffffu, v, (v + i) & 31) + 1; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / N; slot++;
  // 5. LDS dependent chain (pointer chase)
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 7 + 3) & 1023;
  __syncthreads();
  unsigned a = lane;
  t0 = clock64();
  for (int i = 0; i < N; ++i) { a = sm[a]; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / N; slot++;
  v += a;
  // 6. STS -> __syncthreads -> LDS round trip (all warps of the block)
  t0 = clock64();
  for (int i = 0; i < N; ++i) { sm[threadIdx.x] = v; __syncthreads(); v = sm[(threadIdx.x + 32) % blockDim.x] + i; __syncthreads(); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / N; slot++;
  // 7. redux.max + ballot + ffs + 4 shfl (the FPS level-reduce sequence)
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    unsigned m = __reduce_max_sync(0xffffffffu, v ^ lane);
    int src = __ffs(__ballot_sync(0xffffffffu, (v ^ lane) == m)) - 1;
    unsigned a0 = __shfl_sync(0xffffffffu, v, src), a1 = __shfl_sync(0xffffffffu, v + 1, src);
    unsigned a2 = __shfl_sync(0xffffffffu, v + 2, src), a3 = __shfl_sync(0xffffffffu, v + 3, src);
    v = a0 + a1 + a2 + a3 + i;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / N; slot++;
  // 8. same with a shuffle-butterfly arg-max on a packed (key,lane) 64-bit value
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    unsigned long long w = ((unsigned long long)(v ^ lane) << 32) | (31u - lane);
#pragma unroll
    for (int o = 16; o; o >>= 1) { unsigned long long u = __shfl_xor_sync(0xffffffffu, w, o); w = u > w ? u : w; }
    int src = 31 - (int)(w & 31u);
    unsigned a0 = __shfl_sync(0xffffffffu, v, src), a1 = __shfl_sync(0xffffffffu, v + 1, src);
    v = a0 + a1 + i;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / N; slot++;
  // 9. fmnmx dependent chain (ALU latency)
  float f = __uint_as_float(v & 0x3fffffff);
  t0 = clock64();
  for (int i = 0; i < N; ++i) { f = fmaxf(f * 1.0001f, 1.0f); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / N; slot++;
  v += __float_as_uint(f);
  out[threadIdx.x] = v;
}
int main() {
  unsigned *out; long long *cyc;
  cudaMalloc(&out, 4096 * 4); cudaMallocManaged(&cyc, 16 * 8);
  const char *names[] = {"redux.max", "shfl_xor x5 max", "ballot+ffs", "shfl idx", "LDS chase", "STS+bar+LDS+bar",
                         "redux+ballot+4shfl", "bfly argmax64 + 2shfl", "fmul+fmax (2 ops)"};
  for (int threads : {32, 256, 512}) {
    k<<<1, threads>>>(out, cyc, 0);
    cudaDeviceSynchronize();
    printf("threads=%d:", threads);
    for (int i = 0; i < 9; ++i) printf("  %s=%lld", names[i], cyc[i]);
    printf("\n");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
