// Probe for the question "should the ray-attention cores move to tcgen05?" (decide by a probe).
//
// A tensor-core attention keeps, per (query, key, head), the softmax work on the CUDA cores: read the score tile from TMEM, running
// max, exponential, running sum, bf16 (hi, lo) split of P, write P back to TMEM as the A operand of the P.V product.  This file is
// that epilogue, written the way the production kernels write theirs (nfb_tc.cuh primitives: tcgen05.ld.32x32b.x16, packed fp32x2
// arithmetic, cvt.rn.bf16x2, tcgen05.st.32x32b.x8), for one head of one 128-query x 128-key tile, with the MMAs left out -- a LOWER
// bound of what a tensor-core form spends on the CUDA cores.  It is compiled, not run: the number that matters is the SASS instruction
// count of the epilogue per score, to be compared with the CUDA-core attention loop of the production kernel, which spends 28
// instructions per (query, key) for ALL FOUR heads of IBRNet's d_k = 4 attention (profiles/r02j_lines_ray_fwd.txt) and ~20 per
// (query, key, head) in GNT's d = 16 attention.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cubin -o /tmp/probe.cubin tests/probes/probe_tc_attention_epilogue.cu
//   cuobjdump -sass /tmp/probe.cubin | grep -c '^\s*/\*[0-9a-f]\{4\}\*/'        (python profiles/probe_tc_attention.py does both)
#include "../../nerfool_b200/csrc/nfb_tc.cuh"
using namespace nfbtc;

constexpr int KEYS = 128;

// one head: scores in TMEM columns [col_s, col_s + 128), P (bf16 hi | lo) written to [col_p, col_p + 64) | [col_p + 64, col_p + 128)
template <bool SPLIT>
__device__ __forceinline__ void softmax_epilogue(uint32_t tl, int col_s, int col_p, float& l_out) {
  float mx = -3.4e38f;
#pragma unroll
  for (int c = 0; c < KEYS; c += 16) {
    float s[16];
    tmem_ld16(tl + col_s + c, s);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, s[j]);
  }
  const float2 mc = make_float2(-mx * 1.4426950408889634f, -mx * 1.4426950408889634f);
  float2 l2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < KEYS; c += 16) {
    float s[16];
    tmem_ld16(tl + col_s + c, s);
    tmem_ld_wait();
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const float2 t = __ffma2_rn(make_float2(s[j], s[j + 1]), make_float2(1.4426950408889634f, 1.4426950408889634f), mc);
      const float2 p = make_float2(ex2_approx(t.x), ex2_approx(t.y));
      l2 = __fadd2_rn(l2, p);
      if (SPLIT) split_bf16(p.x, p.y, hi[j / 2], lo[j / 2]);
      else hi[j / 2] = pack_bf16(p.x, p.y);
    }
    tmem_st8(tl + col_p + c / 2, hi);
    if (SPLIT) tmem_st8(tl + col_p + 64 + c / 2, lo);
  }
  l_out = l2.x + l2.y;
}

// SPLIT = true: fp32-equivalent P (the default arithmetic of this repo);  false: plain bf16 P
template <bool SPLIT>
__global__ void __launch_bounds__(128) k_probe(uint32_t tmem_base, float* out) {
  const uint32_t tl = tmem_base + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
  float l;
  softmax_epilogue<SPLIT>(tl, 0, 128, l);
  out[blockIdx.x * 128 + threadIdx.x] = l;
}
template __global__ void k_probe<true>(uint32_t, float*);
template __global__ void k_probe<false>(uint32_t, float*);
