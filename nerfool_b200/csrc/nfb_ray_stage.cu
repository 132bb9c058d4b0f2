// IBRNet ray stage (fp32 CUDA-core form): the per-ray half of IBRNet.forward --
// geometry_fc, + pos_encoding, 4-head d_k=4 self-attention over the samples of the ray (row-masked),
// fc + residual + LayerNorm(eps=1e-6), sigma head (mlp_network.py:259-265, 69-119, 23-43) -- and its
// data-gradient.  One CTA per ray (persistent over rays), one thread per sample; K/V (and in the
// backward Q, dO and the softmax statistics) of the ray live in shared memory.
#include "nfb_dense.cuh"
#include "nfb_ray_tc.cuh"
#include "nfb_wgrad_tc.cuh"

namespace {

enum : int {
  R_GEO0 = 0,                   // [65][64]
  RB_GEO0 = R_GEO0 + 65 * 64,   // 64
  R_GEO2 = RB_GEO0 + 64,        // [64][16]
  RB_GEO2 = R_GEO2 + 64 * 16,   // 16
  R_Q = RB_GEO2 + 16,           // [16][16] each, transposed
  R_K = R_Q + 256,
  R_V = R_K + 256,
  R_FC = R_V + 256,
  R_LNW = R_FC + 256,           // 16
  R_LNB = R_LNW + 16,           // 16
  R_OG0 = R_LNB + 16,           // [16][16]
  RB_OG0 = R_OG0 + 256,         // 16
  R_OG2 = RB_OG0 + 16,          // 16 (single output row)
  RB_OG2 = R_OG2 + 16,          // 4
  R_TOTAL = RB_OG2 + 4
};

static __device__ void load_ray_weights(float* sw, const float* __restrict__ p, int tid, int nt) {
  load_wt_transposed(sw + R_GEO0, p + P_GEO0_W, 64, 65, 64, tid, nt);
  load_vec_padded(sw + RB_GEO0, p + P_GEO0_B, 64, 64, tid, nt);
  load_wt_transposed(sw + R_GEO2, p + P_GEO2_W, 16, 64, 16, tid, nt);
  load_vec_padded(sw + RB_GEO2, p + P_GEO2_B, 16, 16, tid, nt);
  load_wt_transposed(sw + R_Q, p + P_ATT_Q, 16, 16, 16, tid, nt);
  load_wt_transposed(sw + R_K, p + P_ATT_K, 16, 16, 16, tid, nt);
  load_wt_transposed(sw + R_V, p + P_ATT_V, 16, 16, 16, tid, nt);
  load_wt_transposed(sw + R_FC, p + P_ATT_FC, 16, 16, 16, tid, nt);
  load_vec_padded(sw + R_LNW, p + P_LN_W, 16, 16, tid, nt);
  load_vec_padded(sw + R_LNB, p + P_LN_B, 16, 16, tid, nt);
  load_wt_transposed(sw + R_OG0, p + P_OG0_W, 16, 16, 16, tid, nt);
  load_vec_padded(sw + RB_OG0, p + P_OG0_B, 16, 16, tid, nt);
  load_vec_padded(sw + R_OG2, p + P_OG2_W, 16, 16, tid, nt);
  load_vec_padded(sw + RB_OG2, p + P_OG2_B, 1, 4, tid, nt);
}

// WG (training) kernel: TMEM accumulator columns of the tensor-core parameter-gradient GEMMs (nfb_wgrad_tc.cuh; the
// reduction runs over the samples of the ray = the threads of the CTA) and the small shared accumulators of the
// vector-shaped tensors (biases, LayerNorm, the single-output layer), which stay exact fp32
enum : int { RC_GEO0 = 0, RC_GEO2 = 64, RC_Q = 80, RC_K = 96, RC_V = 112, RC_FC = 128, RC_OG0 = 144, RC_TOTAL = 160, RC_ALLOC = 256 };
enum : int { RG_GEO0_B = 0, RG_GEO2_B = 64, RG_LNW = 80, RG_LNB = 96, RG_OG0_B = 112, RG_OG2 = 128, RG_OG2_B = 144, RG_TOTAL = 148 };

constexpr float INV_TEMP = 0.5f;   // 1 / sqrt(d_k), d_k = 4 (mlp_network.py:84)
constexpr float LN_EPS = 1e-6f;    // mlp_network.py:87

// geometry_fc on the 65-vector of a sample (streamed from global), + pos-enc  -> hidden (kept for backward)
__device__ __forceinline__ void geometry_fc(const float* sw, const float* __restrict__ psrow, float (&h64)[64],
                                            float (&g16)[16]) {
  load_bias<64>(h64, sw + RB_GEO0);
#pragma unroll
  for (int k4 = 0; k4 < 16; ++k4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(psrow) + k4);
    axpy_row<64>(h64, t.x, sw + R_GEO0 + (4 * k4 + 0) * 64);
    axpy_row<64>(h64, t.y, sw + R_GEO0 + (4 * k4 + 1) * 64);
    axpy_row<64>(h64, t.z, sw + R_GEO0 + (4 * k4 + 2) * 64);
    axpy_row<64>(h64, t.w, sw + R_GEO0 + (4 * k4 + 3) * 64);
  }
  axpy_row<64>(h64, __ldg(psrow + PS_WMEAN), sw + R_GEO0 + 64 * 64);
  elu_inplace<64>(h64);
  load_bias<16>(g16, sw + RB_GEO2);
  dense_acc<64, 16>(sw + R_GEO2, h64, g16);
  elu_inplace<16>(g16);
}

struct AttnOut {
  float o[16];      // normalised attention output, heads concatenated
  float m[4], l[4]; // per-head softmax max / sum (of exp(s - m))
};

// attention for the query row of this thread against all S keys in shared memory
__device__ __forceinline__ void attend(const float (&q)[16], bool row_valid, int S, const float* __restrict__ sk,
                                       const float* __restrict__ sv, AttnOut& r) {
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    float mx = -3.4e38f;
    if (row_valid) {
      for (int j = 0; j < S; ++j) {
        const float4 k = *reinterpret_cast<const float4*>(sk + j * 16 + 4 * h);
        const float s = q[4 * h] * k.x + q[4 * h + 1] * k.y + q[4 * h + 2] * k.z + q[4 * h + 3] * k.w;
        mx = fmaxf(mx, s);
      }
    } else {
      mx = -1e9f;   // masked_fill(mask == 0, -1e9) on the whole query row (mlp_network.py:35-36,105-106)
    }
    float l = 0.f, o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
    for (int j = 0; j < S; ++j) {
      float pj = 1.f;
      if (row_valid) {
        const float4 k = *reinterpret_cast<const float4*>(sk + j * 16 + 4 * h);
        const float s = q[4 * h] * k.x + q[4 * h + 1] * k.y + q[4 * h + 2] * k.z + q[4 * h + 3] * k.w;
        pj = __expf(s - mx);
      }
      const float4 vv = *reinterpret_cast<const float4*>(sv + j * 16 + 4 * h);
      l += pj;
      o0 = fmaf(pj, vv.x, o0); o1 = fmaf(pj, vv.y, o1); o2 = fmaf(pj, vv.z, o2); o3 = fmaf(pj, vv.w, o3);
    }
    const float il = 1.f / l;
    r.o[4 * h] = o0 * il; r.o[4 * h + 1] = o1 * il; r.o[4 * h + 2] = o2 * il; r.o[4 * h + 3] = o3 * il;
    r.m[h] = mx; r.l[h] = l;
  }
}

template <bool BWD, bool WG = false>
__global__ void __launch_bounds__(256, 1) k_ray_stage(int R, int S, const float* __restrict__ ps,
                                                    const float* __restrict__ params,
                                                    const float* __restrict__ pos_enc, float* __restrict__ raw,
                                                    const float* __restrict__ d_raw, float* __restrict__ d_ps,
                                                    float* __restrict__ d_params = nullptr,
                                                    uint8_t* __restrict__ pixel_mask = nullptr) {
  static_assert(!WG || BWD, "parameter gradients are part of the backward");
  extern __shared__ __align__(16) float smem[];
  float* sw = smem;
  float* sk = smem + R_TOTAL;          // [S][16]
  float* sv = sk + (size_t)S * 16;     // [S][16]
  // backward only
  float* sq = sv + (size_t)S * 16;     // [S][16] scaled queries
  float* sdo = sq + (size_t)S * 16;    // [S][16] d(attention output)
  float* sst = sdo + (size_t)S * 16;   // [S][12]: m[4], 1/l[4], D[4]
  float* svalid = sst + (size_t)S * 12;  // [S]
  float* sg = svalid + (((size_t)S + 3) & ~(size_t)3);   // WG: [RG_TOTAL] small accumulators, then the operand tiles
  uint8_t* s_tiles = reinterpret_cast<uint8_t*>(sg + RG_TOTAL);
  const int krows = blockDim.x;                           // rows of a tile = threads of the CTA (multiple of 32)
  uint8_t* s_tb = s_tiles + (size_t)128 * krows * 2;      // B tile [64][krows] bf16 behind the A tile [128][krows]
  uint64_t* s_wbar = reinterpret_cast<uint64_t*>(s_tb + (size_t)64 * krows * 2);
  uint32_t* s_wtmem = reinterpret_cast<uint32_t*>(s_wbar + 1);
  nfbwg::WgTc wg{};
  if (WG) {
    for (int i = threadIdx.x; i < RG_TOTAL; i += blockDim.x) sg[i] = 0.f;
    for (int i = threadIdx.x; i < (128 + 64) * krows / 2; i += blockDim.x) reinterpret_cast<uint32_t*>(s_tiles)[i] = 0u;
    if (threadIdx.x < 32) nfbtc::tmem_alloc(s_wtmem, RC_ALLOC);
    if (threadIdx.x == 0) {
      nfbtc::mbar_init(s_wbar, 1);
      nfbtc::mbar_init_fence();
    }
  }

  load_ray_weights(sw, params, threadIdx.x, blockDim.x);
  if (WG) nfbtc::fence_before_sync();
  __syncthreads();
  if (WG) {
    nfbtc::fence_after_sync();
    wg.sA = s_tiles; wg.sB = s_tb; wg.mbar = s_wbar; wg.tmem = *s_wtmem; wg.phase = 0; wg.pending = false;
    wg.tg = threadIdx.x; wg.bar_id = 1; wg.nthreads = blockDim.x;
    if (threadIdx.x < 128) nfbwg::wg_zero(wg.tmem, threadIdx.x >> 5, RC_TOTAL);
    nfbtc::fence_before_sync();
    __syncthreads();
    nfbtc::fence_after_sync();
  }
  const int s = threadIdx.x;
  const bool act = s < S;

  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    const float* psrow = ps + ((size_t)r * S + (act ? s : 0)) * NFB_PS_STRIDE;
    float h64[64], g16[16];
    geometry_fc(sw, psrow, h64, g16);
    const float nvalid = __ldg(psrow + PS_NVALID);
    const bool row_valid = nvalid > 1.f;
    float xin[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) xin[c] = g16[c] + __ldg(pos_enc + (act ? s : 0) * 16 + c);
    float q[16], kk[16], vv[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { q[c] = 0.f; kk[c] = 0.f; vv[c] = 0.f; }
    dense_acc<16, 16>(sw + R_Q, xin, q);
    dense_acc<16, 16>(sw + R_K, xin, kk);
    dense_acc<16, 16>(sw + R_V, xin, vv);
#pragma unroll
    for (int c = 0; c < 16; ++c) q[c] *= INV_TEMP;
    if (act) {
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
        *reinterpret_cast<float4*>(sk + s * 16 + c) = make_float4(kk[c], kk[c + 1], kk[c + 2], kk[c + 3]);
        *reinterpret_cast<float4*>(sv + s * 16 + c) = make_float4(vv[c], vv[c + 1], vv[c + 2], vv[c + 3]);
      }
    }
    __syncthreads();
    AttnOut at;
    attend(q, row_valid, S, sk, sv, at);

    // fc + residual + LayerNorm
    float y[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) y[c] = xin[c];
    dense_acc<16, 16>(sw + R_FC, at.o, y);
    float mu = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) mu += y[c];
    mu *= (1.f / 16.f);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) var = fmaf(y[c] - mu, y[c] - mu, var);
    var *= (1.f / 16.f);
    const float rstd = rsqrtf(var + LN_EPS);
    float xhat[16], ln[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      xhat[c] = (y[c] - mu) * rstd;
      ln[c] = fmaf(xhat[c], sw[R_LNW + c], sw[R_LNB + c]);
    }
    // sigma head
    float hh[16];
    load_bias<16>(hh, sw + RB_OG0);
    dense_acc<16, 16>(sw + R_OG0, ln, hh);
    elu_inplace<16>(hh);
    const float z2 = dot_row<16>(hh, sw + R_OG2) + sw[RB_OG2];

    if (!BWD) {
      float sigma = fmaxf(z2, 0.f);
      if (nvalid < 1.f) sigma = 0.f;                     // mlp_network.py:265
      if (act) {
        reinterpret_cast<float4*>(raw)[(size_t)r * S + s] =
            make_float4(__ldg(psrow + PS_RGB), __ldg(psrow + PS_RGB + 1), __ldg(psrow + PS_RGB + 2), sigma);
        if (pixel_mask) pixel_mask[(size_t)r * S + s] = nvalid > 1.f ? 1 : 0;
      }
      __syncthreads();                                   // sk / sv are rewritten by the next ray
      continue;
    }

    // =================================== backward ===================================
    if (BWD) {
      float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
      if (act) dr = __ldg(reinterpret_cast<const float4*>(d_raw) + (size_t)r * S + s);
      const float dz2 = (z2 > 0.f && !(nvalid < 1.f)) ? dr.w : 0.f;
      const float gate = act ? 1.f : 0.f;
      // sigma head
      float dln[16];
      {
        float dh[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) dh[k] = dz2 * sw[R_OG2 + k] * elu_grad_from_out(hh[k]);
        dense_T<16, 16>(sw + R_OG0, dh, dln);
        if (WG) {
          wgrad_rowsum<16>(sg + RG_OG2, hh, dz2 * gate);
          const float dbz[4] = {dz2, 0.f, 0.f, 0.f};
          wgrad_rowsum<4>(sg + RG_OG2_B, dbz, gate);
          nfbwg::wgrad_tc<16, 16>(wg, RC_OG0, ln, dh, gate);
          wgrad_rowsum<16>(sg + RG_OG0_B, dh, gate);
          float t[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) t[c] = dln[c] * xhat[c];
          wgrad_rowsum<16>(sg + RG_LNW, t, gate);
          wgrad_rowsum<16>(sg + RG_LNB, dln, gate);
        }
      }
      // LayerNorm backward: dy = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dln * gamma
      float dy[16];
      {
        float gsum = 0.f, gx = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          dy[c] = dln[c] * sw[R_LNW + c];
          gsum += dy[c];
          gx = fmaf(dy[c], xhat[c], gx);
        }
        gsum *= (1.f / 16.f); gx *= (1.f / 16.f);
#pragma unroll
        for (int c = 0; c < 16; ++c) dy[c] = rstd * (dy[c] - gsum - xhat[c] * gx);
      }
      // y = fc(o) + xin
      float dO[16];
      dense_T<16, 16>(sw + R_FC, dy, dO);
      if (WG) nfbwg::wgrad_tc<16, 16>(wg, RC_FC, at.o, dy, gate);
      // publish per-query quantities for the key-side pass
      if (act) {
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          *reinterpret_cast<float4*>(sq + s * 16 + c) = make_float4(q[c], q[c + 1], q[c + 2], q[c + 3]);
          *reinterpret_cast<float4*>(sdo + s * 16 + c) = make_float4(dO[c], dO[c + 1], dO[c + 2], dO[c + 3]);
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          sst[s * 12 + h] = at.m[h];
          sst[s * 12 + 4 + h] = 1.f / at.l[h];
          sst[s * 12 + 8 + h] = dO[4 * h] * at.o[4 * h] + dO[4 * h + 1] * at.o[4 * h + 1] +
                                dO[4 * h + 2] * at.o[4 * h + 2] + dO[4 * h + 3] * at.o[4 * h + 3];
        }
        svalid[s] = row_valid ? 1.f : 0.f;
      }
      __syncthreads();
      // query side: dq_i = sum_j dS_ij k_j  (zero for masked rows: masked_fill blocks the gradient)
      float dq[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) dq[c] = 0.f;
      if (row_valid) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float il = 1.f / at.l[h];
          const float Dh = dO[4 * h] * at.o[4 * h] + dO[4 * h + 1] * at.o[4 * h + 1] +
                           dO[4 * h + 2] * at.o[4 * h + 2] + dO[4 * h + 3] * at.o[4 * h + 3];
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          for (int j = 0; j < S; ++j) {
            const float4 k = *reinterpret_cast<const float4*>(sk + j * 16 + 4 * h);
            const float4 vj = *reinterpret_cast<const float4*>(sv + j * 16 + 4 * h);
            const float sc = q[4 * h] * k.x + q[4 * h + 1] * k.y + q[4 * h + 2] * k.z + q[4 * h + 3] * k.w;
            const float pj = __expf(sc - at.m[h]) * il;
            const float dP = dO[4 * h] * vj.x + dO[4 * h + 1] * vj.y + dO[4 * h + 2] * vj.z + dO[4 * h + 3] * vj.w;
            const float dS = pj * (dP - Dh);
            a0 = fmaf(dS, k.x, a0); a1 = fmaf(dS, k.y, a1); a2 = fmaf(dS, k.z, a2); a3 = fmaf(dS, k.w, a3);
          }
          dq[4 * h] = a0 * INV_TEMP; dq[4 * h + 1] = a1 * INV_TEMP; dq[4 * h + 2] = a2 * INV_TEMP; dq[4 * h + 3] = a3 * INV_TEMP;
        }
      }
      // key side: dk_j = sum_i dS_ij q_i ; dv_j = sum_i p_ij dO_i   (this thread is key j = s)
      float dk[16], dv[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
      {
        const float invS = 1.f / (float)S;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float k0 = kk[4 * h], k1 = kk[4 * h + 1], k2 = kk[4 * h + 2], k3 = kk[4 * h + 3];
          const float v0 = vv[4 * h], v1 = vv[4 * h + 1], v2 = vv[4 * h + 2], v3 = vv[4 * h + 3];
          float ak0 = 0.f, ak1 = 0.f, ak2 = 0.f, ak3 = 0.f, av0 = 0.f, av1 = 0.f, av2 = 0.f, av3 = 0.f;
          for (int i = 0; i < S; ++i) {
            const float4 qi = *reinterpret_cast<const float4*>(sq + i * 16 + 4 * h);
            const float4 di = *reinterpret_cast<const float4*>(sdo + i * 16 + 4 * h);
            const bool vi = svalid[i] != 0.f;
            float pij, dS = 0.f;
            if (vi) {
              const float sc = qi.x * k0 + qi.y * k1 + qi.z * k2 + qi.w * k3;
              pij = __expf(sc - sst[i * 12 + h]) * sst[i * 12 + 4 + h];
              const float dP = di.x * v0 + di.y * v1 + di.z * v2 + di.w * v3;
              dS = pij * (dP - sst[i * 12 + 8 + h]);
            } else {
              pij = invS;
            }
            ak0 = fmaf(dS, qi.x, ak0); ak1 = fmaf(dS, qi.y, ak1); ak2 = fmaf(dS, qi.z, ak2); ak3 = fmaf(dS, qi.w, ak3);
            av0 = fmaf(pij, di.x, av0); av1 = fmaf(pij, di.y, av1); av2 = fmaf(pij, di.z, av2); av3 = fmaf(pij, di.w, av3);
          }
          dk[4 * h] = ak0; dk[4 * h + 1] = ak1; dk[4 * h + 2] = ak2; dk[4 * h + 3] = ak3;
          dv[4 * h] = av0; dv[4 * h + 1] = av1; dv[4 * h + 2] = av2; dv[4 * h + 3] = av3;
        }
      }
      if (WG) {
        nfbwg::wgrad_tc<16, 16>(wg, RC_Q, xin, dq, gate);
        nfbwg::wgrad_tc<16, 16>(wg, RC_K, xin, dk, gate);
        nfbwg::wgrad_tc<16, 16>(wg, RC_V, xin, dv, gate);
      }
      // d xin = dy (residual) + Wq^T dq + Wk^T dk + Wv^T dv ; pos_encoding is a constant
      float dx[16];
      {
        float t[16];
        dense_T<16, 16>(sw + R_Q, dq, t);
#pragma unroll
        for (int c = 0; c < 16; ++c) dx[c] = dy[c] + t[c];
        dense_T<16, 16>(sw + R_K, dk, t);
#pragma unroll
        for (int c = 0; c < 16; ++c) dx[c] += t[c];
        dense_T<16, 16>(sw + R_V, dv, t);
#pragma unroll
        for (int c = 0; c < 16; ++c) dx[c] += t[c];
      }
      // geometry_fc backward
#pragma unroll
      for (int c = 0; c < 16; ++c) dx[c] *= elu_grad_from_out(g16[c]);
      float dh64[64];
      dense_T<64, 16>(sw + R_GEO2, dx, dh64);
#pragma unroll
      for (int k = 0; k < 64; ++k) dh64[k] *= elu_grad_from_out(h64[k]);
      if (WG) {
        nfbwg::wgrad_tc<64, 16>(wg, RC_GEO2, h64, dx, gate);
        wgrad_rowsum<16>(sg + RG_GEO2_B, dx, gate);
        float xin65[65];
#pragma unroll
        for (int k = 0; k < 65; ++k) xin65[k] = __ldg(psrow + k);
        nfbwg::wgrad_tc<65, 64>(wg, RC_GEO0, xin65, dh64, gate);
        wgrad_rowsum<64>(sg + RG_GEO0_B, dh64, gate);
      }
      if (act) {
        float* out = d_ps + ((size_t)r * S + s) * NFB_PS_STRIDE;
#pragma unroll
        for (int k4 = 0; k4 < 16; ++k4) {
          float4 t;
          t.x = dot_row<64>(dh64, sw + R_GEO0 + (4 * k4 + 0) * 64);
          t.y = dot_row<64>(dh64, sw + R_GEO0 + (4 * k4 + 1) * 64);
          t.z = dot_row<64>(dh64, sw + R_GEO0 + (4 * k4 + 2) * 64);
          t.w = dot_row<64>(dh64, sw + R_GEO0 + (4 * k4 + 3) * 64);
          reinterpret_cast<float4*>(out)[k4] = t;
        }
        const float dwm = dot_row<64>(dh64, sw + R_GEO0 + 64 * 64);
        reinterpret_cast<float4*>(out)[16] = make_float4(dwm, dr.x, dr.y, dr.z);
        reinterpret_cast<float4*>(out)[17] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
    }
  }
  if (WG) {
    nfbwg::wg_wait(wg);
    nfbtc::fence_before_sync();
    __syncthreads();
    nfbtc::fence_after_sync();
    float* dp = d_params;
    if (threadIdx.x < 128) {
      const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
      nfbwg::wg_flush(wg.tmem, w, l, RC_GEO0, 65, 64, dp + P_GEO0_W);
      nfbwg::wg_flush(wg.tmem, w, l, RC_GEO2, 64, 16, dp + P_GEO2_W);
      nfbwg::wg_flush(wg.tmem, w, l, RC_Q, 16, 16, dp + P_ATT_Q);
      nfbwg::wg_flush(wg.tmem, w, l, RC_K, 16, 16, dp + P_ATT_K);
      nfbwg::wg_flush(wg.tmem, w, l, RC_V, 16, 16, dp + P_ATT_V);
      nfbwg::wg_flush(wg.tmem, w, l, RC_FC, 16, 16, dp + P_ATT_FC);
      nfbwg::wg_flush(wg.tmem, w, l, RC_OG0, 16, 16, dp + P_OG0_W);
    }
    flush_vec(dp + P_GEO0_B, sg + RG_GEO0_B, 64, threadIdx.x, blockDim.x);
    flush_vec(dp + P_GEO2_B, sg + RG_GEO2_B, 16, threadIdx.x, blockDim.x);
    flush_vec(dp + P_LN_W, sg + RG_LNW, 16, threadIdx.x, blockDim.x);
    flush_vec(dp + P_LN_B, sg + RG_LNB, 16, threadIdx.x, blockDim.x);
    flush_vec(dp + P_OG0_B, sg + RG_OG0_B, 16, threadIdx.x, blockDim.x);
    flush_vec(dp + P_OG2_W, sg + RG_OG2, 16, threadIdx.x, blockDim.x);
    flush_vec(dp + P_OG2_B, sg + RG_OG2_B, 1, threadIdx.x, blockDim.x);
    nfbtc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) nfbtc::tmem_dealloc(wg.tmem, RC_ALLOC);
  }
}

inline int ray_block(int S) { return ((S + 31) / 32) * 32; }

}  // namespace

extern "C" size_t nfb_ray_stash_bytes(int R, int S) {
  if (R <= 0 || S < 1 || S > NFB_MAX_SAMPLES) return 0;
  if (S > nfbrtc::GROUP) return (size_t)R * 2 * nfbrtc::RP_TILE_BYTES;     // two 128-row tiles per ray
  const int rpg = nfbrtc::GROUP / S;
  return (size_t)((R + rpg - 1) / rpg) * nfbrtc::RP_TILE_BYTES;
}

extern "C" int nfb_ibrnet_ray_fwd(int R, int S, const float* ps, const float* params, const float* pos_enc,
                                  float* raw, uint8_t* pixel_mask, float* stash, int precision, void* stream) {
  NFB_REQUIRE(R >= 0 && S >= 1, NFB_EINVAL, "nfb_ibrnet_ray_fwd: bad arguments (R=%d S=%d)", R, S);
  NFB_REQUIRE(S <= NFB_MAX_SAMPLES, NFB_EUNSUPPORTED, "nfb_ibrnet_ray_fwd: S=%d > %d samples per ray", S, NFB_MAX_SAMPLES);
  if (R == 0) return NFB_OK;
  NFB_REQUIRE(ps && params && pos_enc && raw, NFB_EINVAL, "nfb_ibrnet_ray_fwd: NULL buffer");
  NFB_REQUIRE(((uintptr_t)ps % 16) == 0 && ((uintptr_t)raw % 16) == 0, NFB_EINVAL, "nfb_ibrnet_ray_fwd: ps/raw must be 16-byte aligned");
  NFB_REQUIRE(precision >= NFB_PREC_FP32 && precision <= NFB_PREC_BF16, NFB_EINVAL, "nfb_ibrnet_ray_fwd: bad precision %d", precision);
  if (stash)
    NFB_REQUIRE(precision != NFB_PREC_FP32 && ((uintptr_t)stash % 16) == 0, NFB_EUNSUPPORTED,
                "nfb_ibrnet_ray_fwd: the activation stash exists for the tensor-core forms only, 16-byte aligned");
  if (precision != NFB_PREC_FP32) {
    nfbrtc::RayArgs a{R, S, ps, params, pos_enc, raw, nullptr, nullptr, stash, pixel_mask};
    cudaStream_t st = (cudaStream_t)stream;
    if (stash) return precision == NFB_PREC_BF16 ? nfb_launch_ray_tc_fwd_p1_save(a, st) : nfb_launch_ray_tc_fwd_p3_save(a, st);
    return precision == NFB_PREC_BF16 ? nfb_launch_ray_tc_fwd_p1(a, st) : nfb_launch_ray_tc_fwd_p3(a, st);
  }
  const size_t smem = (size_t)(R_TOTAL + 2 * S * 16) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(k_ray_stage<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "nfb_ibrnet_ray_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int block = ray_block(S);
  int per_sm = 2048 / block; if (per_sm > 4) per_sm = 4;
  int grid = nfb_num_sms() * per_sm; if (grid > R) grid = R;
  k_ray_stage<false><<<grid, block, smem, (cudaStream_t)stream>>>(R, S, ps, params, pos_enc, raw, nullptr, nullptr, nullptr, pixel_mask);
  NFB_CHECK_LAUNCH("k_ray_stage<fwd>");
  return NFB_OK;
}

extern "C" int nfb_ibrnet_ray_bwd(int R, int S, const float* ps, const float* params, const float* pos_enc,
                                  const float* d_raw, float* d_ps, const float* stash, int precision, void* stream) {
  NFB_REQUIRE(R >= 0 && S >= 1, NFB_EINVAL, "nfb_ibrnet_ray_bwd: bad arguments (R=%d S=%d)", R, S);
  NFB_REQUIRE(S <= NFB_MAX_SAMPLES, NFB_EUNSUPPORTED, "nfb_ibrnet_ray_bwd: S=%d > %d samples per ray", S, NFB_MAX_SAMPLES);
  if (R == 0) return NFB_OK;
  NFB_REQUIRE(ps && params && pos_enc && d_raw && d_ps, NFB_EINVAL, "nfb_ibrnet_ray_bwd: NULL buffer");
  NFB_REQUIRE(((uintptr_t)ps % 16) == 0 && ((uintptr_t)d_raw % 16) == 0 && ((uintptr_t)d_ps % 16) == 0, NFB_EINVAL,
              "nfb_ibrnet_ray_bwd: ps/d_raw/d_ps must be 16-byte aligned");
  NFB_REQUIRE(precision >= NFB_PREC_FP32 && precision <= NFB_PREC_BF16, NFB_EINVAL, "nfb_ibrnet_ray_bwd: bad precision %d", precision);
  if (stash)
    NFB_REQUIRE(precision != NFB_PREC_FP32 && ((uintptr_t)stash % 16) == 0, NFB_EUNSUPPORTED,
                "nfb_ibrnet_ray_bwd: the activation stash exists for the tensor-core forms only, 16-byte aligned");
  if (precision != NFB_PREC_FP32) {
    nfbrtc::RayArgs a{R, S, ps, params, pos_enc, nullptr, d_raw, d_ps, const_cast<float*>(stash)};
    cudaStream_t st = (cudaStream_t)stream;
    if (stash) return precision == NFB_PREC_BF16 ? nfb_launch_ray_tc_bwd_stash_p1(a, st) : nfb_launch_ray_tc_bwd_stash_p3(a, st);
    return precision == NFB_PREC_BF16 ? nfb_launch_ray_tc_bwd_p1(a, st) : nfb_launch_ray_tc_bwd_p3(a, st);
  }
  const size_t smem = (size_t)(R_TOTAL + 4 * S * 16 + S * 12 + S) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(k_ray_stage<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "nfb_ibrnet_ray_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int block = ray_block(S);
  int per_sm = 2048 / block; if (per_sm > 2) per_sm = 2;
  int grid = nfb_num_sms() * per_sm; if (grid > R) grid = R;
  k_ray_stage<true><<<grid, block, smem, (cudaStream_t)stream>>>(R, S, ps, params, pos_enc, nullptr, d_raw, d_ps);
  NFB_CHECK_LAUNCH("k_ray_stage<bwd>");
  return NFB_OK;
}

extern "C" int nfb_ibrnet_ray_wgrad(int R, int S, const float* ps, const float* params, const float* pos_enc,
                                    const float* d_raw, float* d_ps, float* d_params, void* stream) {
  NFB_REQUIRE(R >= 0 && S >= 1, NFB_EINVAL, "nfb_ibrnet_ray_wgrad: bad arguments (R=%d S=%d)", R, S);
  NFB_REQUIRE(S <= NFB_MAX_SAMPLES, NFB_EUNSUPPORTED, "nfb_ibrnet_ray_wgrad: S=%d > %d samples per ray", S, NFB_MAX_SAMPLES);
  if (R == 0) return NFB_OK;
  NFB_REQUIRE(ps && params && pos_enc && d_raw && d_ps && d_params, NFB_EINVAL, "nfb_ibrnet_ray_wgrad: NULL buffer");
  NFB_REQUIRE(((uintptr_t)ps % 16) == 0 && ((uintptr_t)d_raw % 16) == 0 && ((uintptr_t)d_ps % 16) == 0, NFB_EINVAL,
              "nfb_ibrnet_ray_wgrad: ps/d_raw/d_ps must be 16-byte aligned");
  const int wblock = ray_block(S) < 128 ? 128 : ray_block(S);     // >= 4 warps: they own the 128 TMEM lanes of the accumulators
  const size_t smem = (size_t)(R_TOTAL + 4 * S * 16 + S * 12 + ((S + 3) & ~3) + RG_TOTAL) * sizeof(float) +
                      (size_t)(128 + 64) * wblock * 2 + 8 + 16;
  cudaError_t e = cudaFuncSetAttribute(k_ray_stage<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "nfb_ibrnet_ray_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int block = wblock;
  int per_sm = 2;                                                  // 2 x 256 TMEM columns per SM
  int grid = nfb_num_sms() * per_sm; if (grid > R) grid = R;
  k_ray_stage<true, true><<<grid, block, smem, (cudaStream_t)stream>>>(R, S, ps, params, pos_enc, nullptr, d_raw, d_ps, d_params);
  NFB_CHECK_LAUNCH("k_ray_stage<bwd,wgrad>");
  return NFB_OK;
}
