// Thread-per-row dense layers for the fp32 (CUDA-core) IBRNet kernels.
// Every thread owns one row of activations in registers; the layer weights sit in shared memory,
// TRANSPOSED to [K][NP] (NP = N rounded up to 4, padding columns zero) so that one broadcast LDS.128
// feeds four FMAs in the forward direction and four terms of a dot product in the backward direction.
#pragma once
#include "nfb_common.cuh"

// out[0..NP) += x * Wt_row[0..NP)
// The fused multiply-adds are issued as packed fp32x2 (FFMA2: two independent round-to-nearest FMAs per instruction, so the results
// are bit-identical to four scalar FFMAs at half the issue slots; measured on the GNT reverse sweep: 35.3 -> 30.1 ms).
// -DNFB_DENSE_SCALAR restores the scalar form.
template <int NP>
__device__ __forceinline__ void axpy_row(float (&out)[NP], float x, const float* __restrict__ wrow) {
#pragma unroll
  for (int n = 0; n < NP; n += 4) {
    const float4 w = *reinterpret_cast<const float4*>(wrow + n);
#ifndef NFB_DENSE_SCALAR
    const float2 xx = make_float2(x, x);
    const float2 r0 = __ffma2_rn(xx, make_float2(w.x, w.y), make_float2(out[n + 0], out[n + 1]));
    const float2 r1 = __ffma2_rn(xx, make_float2(w.z, w.w), make_float2(out[n + 2], out[n + 3]));
    out[n + 0] = r0.x; out[n + 1] = r0.y; out[n + 2] = r1.x; out[n + 3] = r1.y;
#else
    out[n + 0] = fmaf(x, w.x, out[n + 0]);
    out[n + 1] = fmaf(x, w.y, out[n + 1]);
    out[n + 2] = fmaf(x, w.z, out[n + 2]);
    out[n + 3] = fmaf(x, w.w, out[n + 3]);
#endif
  }
}

template <int NP>
__device__ __forceinline__ void load_bias(float (&out)[NP], const float* __restrict__ b) {
#pragma unroll
  for (int n = 0; n < NP; n += 4) {
    const float4 w = *reinterpret_cast<const float4*>(b + n);
    out[n + 0] = w.x; out[n + 1] = w.y; out[n + 2] = w.z; out[n + 3] = w.w;
  }
}

// out[NP] += in[K] . Wt[K][NP]
template <int K, int NP>
__device__ __forceinline__ void dense_acc(const float* __restrict__ Wt, const float (&in)[K], float (&out)[NP]) {
#pragma unroll
  for (int k = 0; k < K; ++k) axpy_row<NP>(out, in[k], Wt + k * NP);
}

// sum_n dy[n] * wrow[n]
template <int NP>
__device__ __forceinline__ float dot_row(const float (&dy)[NP], const float* __restrict__ wrow) {
#ifndef NFB_DENSE_SCALAR
  float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
#pragma unroll
  for (int n = 0; n < NP; n += 4) {
    const float4 w = *reinterpret_cast<const float4*>(wrow + n);
    a01 = __ffma2_rn(make_float2(dy[n + 0], dy[n + 1]), make_float2(w.x, w.y), a01);
    a23 = __ffma2_rn(make_float2(dy[n + 2], dy[n + 3]), make_float2(w.z, w.w), a23);
  }
  return (a01.x + a01.y) + (a23.x + a23.y);
#else
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int n = 0; n < NP; n += 4) {
    const float4 w = *reinterpret_cast<const float4*>(wrow + n);
    a0 = fmaf(dy[n + 0], w.x, a0);
    a1 = fmaf(dy[n + 1], w.y, a1);
    a2 = fmaf(dy[n + 2], w.z, a2);
    a3 = fmaf(dy[n + 3], w.w, a3);
  }
  return (a0 + a1) + (a2 + a3);
#endif
}

// dx[k] = sum_n dy[n] * Wt[k][n]   (transpose product with the same smem layout)
template <int K, int NP>
__device__ __forceinline__ void dense_T(const float* __restrict__ Wt, const float (&dy)[NP], float (&dx)[K]) {
#pragma unroll
  for (int k = 0; k < K; ++k) dx[k] = dot_row<NP>(dy, Wt + k * NP);
}

template <int N>
__device__ __forceinline__ void elu_inplace(float (&a)[N]) {
#pragma unroll
  for (int n = 0; n < N; ++n) a[n] = elu_f(a[n]);
}

// cooperative load of a torch-layout weight [N][K] (global) into transposed smem [K][NP]
__device__ __forceinline__ void load_wt_transposed(float* __restrict__ dst, const float* __restrict__ src, int N,
                                                   int K, int NP, int tid, int nthreads) {
  for (int i = tid; i < K * NP; i += nthreads) {
    const int k = i / NP, n = i - k * NP;
    dst[i] = (n < N) ? __ldg(src + n * K + k) : 0.f;
  }
}
__device__ __forceinline__ void load_vec_padded(float* __restrict__ dst, const float* __restrict__ src, int N, int NP,
                                                int tid, int nthreads) {
  for (int i = tid; i < NP; i += nthreads) dst[i] = (i < N) ? __ldg(src + i) : 0.f;
}

// ---------------------------------------------------------------------------------------------------
// parameter gradients (training).  The weight gradients are GEMMs over the row index on the tensor cores
// (nfb_wgrad_tc.cuh); vector-shaped gradients (biases, LayerNorm, single-output layers) are plain row sums:
// 32 values per lane are reduced across the warp with a 5-stage exchange butterfly (31 shuffles per 32 values, lane
// L ends up with the total of element n0 + L) and added with one shared-memory atomic.
// ---------------------------------------------------------------------------------------------------
// sg[n] += scale * (sum over the warp's rows of dy[n]), fully inlined with compile-time register indices (the caller's
// array stays in registers): the same exchange butterfly, 31 shuffles per 32 values
template <int NP>
__device__ __forceinline__ void wgrad_rowsum(float* __restrict__ sg, const float (&dy)[NP], float scale) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int n0 = 0; n0 < NP; n0 += 32) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (n0 + i < NP) ? dy[(n0 + i < NP) ? n0 + i : 0] * scale : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < o; ++i) {
        const float keep = up ? v[i + o] : v[i];
        const float send = up ? v[i] : v[i + o];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    if (n0 + lane < NP) atomicAdd(sg + n0 + lane, v[0]);
  }
}

// add a transposed [K][NP] shared-memory accumulator into the torch-layout [N][K] gradient blob
__device__ __forceinline__ void flush_wt_transposed(float* __restrict__ dst, const float* __restrict__ sg, int N, int K,
                                                    int NP, int tid, int nthreads) {
  for (int i = tid; i < K * NP; i += nthreads) {
    const int k = i / NP, n = i - k * NP;
    if (n < N) atomicAdd(dst + n * K + k, sg[i]);
  }
}
__device__ __forceinline__ void flush_vec(float* __restrict__ dst, const float* __restrict__ sg, int N, int tid, int nthreads) {
  for (int i = tid; i < N; i += nthreads) atomicAdd(dst + i, sg[i]);
}
