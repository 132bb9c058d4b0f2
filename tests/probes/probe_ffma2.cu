// Micro-probe: issue throughput of scalar FFMA against packed FFMA2 (fma.rn.f32x2, sm_100a) with 8 independent chains
// per thread.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_ffma2 tests/probes/probe_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float a, float b) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a2.x, b2.x); acc[i].y = fmaf(acc[i].y, a2.y, b2.y); }
        else acc[i] = __ffma2_rn(acc[i], a2, b2);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 4 * 512 * 4);
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode) {
    for (int blocks_per_sm = 1; blocks_per_sm <= 4; blocks_per_sm *= 2) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * blocks_per_sm, 512>>>(d, iters, 0.999f, 0.001f);
        else k<1><<<148 * blocks_per_sm, 512>>>(d, iters, 0.999f, 0.001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)148 * blocks_per_sm * 512 * iters * 4 * 8 * 2;
      printf("%s blocks/SM=%d: %.3f ms  %.1f TFMA/s (%.1f TFLOP/s)\n", mode ? "FFMA2" : "FFMA ", blocks_per_sm, ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
    }
  }
  return 0;
}
