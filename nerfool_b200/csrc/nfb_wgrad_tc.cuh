// Parameter gradients on the tensor cores: dW[out][in] = sum over rows of dY[row][out] * X[row][in] is a GEMM whose
// reduction dimension is the ROW index, so the two per-row vectors are staged (as bf16) into shared memory in the
// canonical K-major operand layout with K = the 128 rows of the group's tile,
//     A = X^T  (M = input index),   B = dY^T (N = output index),
// and one tcgen05.mma.ss batch (8 K-steps of 16 rows) adds the tile's contribution to the layer's fp32 accumulator
// D[input index (TMEM lane)][output index (TMEM column)], which stays resident in TMEM for the whole kernel and is
// flushed to global memory once per CTA at the end.  Operands are rounded to bf16 (products exact, fp32 accumulation);
// the data-gradient path of the same kernel and the bias gradients (plain row sums, warp butterfly) stay fp32.
#pragma once
#include "nfb_tc.cuh"

namespace nfbwg {
using namespace nfbtc;

constexpr int A_TILE_BYTES = 128 * 128 * 2;     // [M = 128][K = 128 rows] bf16
constexpr int B_TILE_BYTES = 64 * 128 * 2;      // [N <= 64][K = 128 rows] bf16

struct WgTc {
  uint8_t* sA;
  uint8_t* sB;
  uint64_t* mbar;
  uint32_t tmem;       // TMEM base address (lane 0, column 0) of the CTA's allocation
  uint32_t phase;      // parity the next completion will have
  bool pending;        // an MMA batch of this group is still reading the staging tiles
  int tg, bar_id;
  int nthreads;        // threads (= rows = MMA K extent, a multiple of 16) of the group; 0 means 128
};

__device__ __forceinline__ void st_bf16(uint8_t* p, float v) { *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(v); }

__device__ __forceinline__ void wg_wait(WgTc& c) {
  if (c.pending) {
    mbar_wait(c.mbar, c.phase);
    c.phase ^= 1u;
    c.pending = false;
  }
}

// accumulate this tile's outer products of layer (K_IN inputs, N_ARR outputs incl. padding) into TMEM columns [dcol, dcol + NP)
template <int K_IN, int N_ARR>
__device__ __forceinline__ void wgrad_tc(WgTc& c, int dcol, const float (&x)[K_IN], const float (&dy)[N_ARR], float gate) {
  constexpr int NP = (N_ARR + 15) / 16 * 16;
  static_assert(K_IN <= 128 && NP <= 64, "tile sizes");
  wg_wait(c);                                                   // the previous batch has finished reading the tiles
  const int k = c.tg;
  uint8_t* pa = c.sA + (k >> 3) * 2048 + (k & 7) * 2;           // element (m, k): (k/8)*(128*16) + (m/8)*128 + (m%8)*16 + (k%8)*2
#pragma unroll
  for (int m = 0; m < K_IN; ++m) st_bf16(pa + (m >> 3) * 128 + (m & 7) * 16, x[m] * gate);
  uint8_t* pb = c.sB + (k >> 3) * (NP * 16) + (k & 7) * 2;      // element (n, k): (k/8)*(NP*16) + (n/8)*128 + (n%8)*16 + (k%8)*2
#pragma unroll
  for (int n = 0; n < NP; ++n) st_bf16(pb + (n >> 3) * 128 + (n & 7) * 16, n < N_ARR ? dy[n < N_ARR ? n : 0] * gate : 0.f);
  fence_proxy_async_smem();                                     // generic-proxy stores -> visible to the MMA (async proxy)
  const int rows = c.nthreads ? c.nthreads : 128;
  named_bar_sync(c.bar_id, rows);
  if (c.tg == 0) {
    fence_after_sync();
    constexpr uint32_t idesc = idesc_bf16(128, NP);
    const uint32_t a_addr = smem_u32(c.sA), b_addr = smem_u32(c.sB);
    for (int ks = 0; ks < rows / 16; ++ks)
      mma_ss(c.tmem + dcol, smem_desc(a_addr + ks * 4096, 2048, 128), smem_desc(b_addr + ks * 2 * NP * 16, NP * 16, 128), idesc, true);
    mma_commit(c.mbar);
  }
  c.pending = true;
}

// zero the accumulator columns [0, ncols) (ncols multiple of 16); called by the 4 warps that own TMEM lanes 0..127
__device__ __forceinline__ void wg_zero(uint32_t tmem, int warp, int ncols) {
  const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t z[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) z[i] = 0u;
  for (int c0 = 0; c0 < ncols; c0 += 16) tmem_st16(tl + c0, z);
  tmem_st_wait();
}

// add accumulator region (lane = input index m, column dcol + n) into dW[n][m] (torch layout [out][in])
__device__ __forceinline__ void wg_flush(uint32_t tmem, int warp, int lane, int dcol, int n_in, int n_out, float* __restrict__ dW) {
  const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int m = (warp & 3) * 32 + lane;
  for (int n0 = 0; n0 < n_out; n0 += 16) {
    float v[16];
    tmem_ld16(tl + dcol + n0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int n = n0 + j;
      if (n < n_out) {
        if (m < n_in) atomicAdd(dW + (size_t)n * n_in + m, v[j]);
      }
    }
  }
}

}  // namespace nfbwg
