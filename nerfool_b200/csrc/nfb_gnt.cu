// GNT forward (gnt/transformer_network.py:205-309): view transformer (subtraction attention over the source
// views, softmax per channel) + ray transformer (4-head self-attention over the samples of a ray), netwidth 64.
// First CUDA form of SURVEY 8 row a15 / f3: fp32 on the CUDA cores, one thread per sample, layer weights resident
// in shared memory (transposed, nfb_dense.cuh), activations q[N][64] and the projected view features F[N][V][64]
// in a caller-provided workspace.  Forward only (the render path of BASELINE config 5); the tcgen05 form and the
// data-gradient are the next step.
//
// Launch sequence of nfb_gnt_fwd: embed (rgbfeat_fc + max over views) -> per layer { view attention, FFN,
// [q_fc with positional encodings on even layers], ray attention, FFN } -> head (LayerNorm, mean over samples,
// rgb_fc).  Dropout is the identity (eval mode, transformer_network.py:45,72,136).
#include "nfb_gnt_common.cuh"
#include "nfb_gnt_tc.cuh"

using namespace nfbgnt;

namespace {

// ---------------------------------------------------------------------------------------------------
// embed: F[row] = rgbfeat_fc(rgb_feat[row])  (35 -> 64 ReLU -> 64), one thread per (sample, view) row
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_gnt_embed(size_t rows, const float* __restrict__ rgb_feat,
                                                    const float* __restrict__ params, float* __restrict__ F) {
  extern __shared__ __align__(16) float sm[];
  float* w0 = sm;                 // [35][64]
  float* b0 = w0 + 35 * D;
  float* w2 = b0 + D;             // [64][64]
  float* b2 = w2 + D * D;
  load_wt_transposed(w0, params + G_RF0_W, D, 35, D, threadIdx.x, blockDim.x);
  load_vec_padded(b0, params + G_RF0_B, D, D, threadIdx.x, blockDim.x);
  load_wt_transposed(w2, params + G_RF2_W, D, D, D, threadIdx.x, blockDim.x);
  load_vec_padded(b2, params + G_RF2_B, D, D, threadIdx.x, blockDim.x);
  __syncthreads();
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
    float h[D];
    load_bias<D>(h, b0);
    const float* x = rgb_feat + r * NFB_ROW_CH;
#pragma unroll
    for (int k = 0; k < NFB_ROW_CH; ++k) axpy_row<D>(h, __ldg(x + k), w0 + k * D);
#pragma unroll
    for (int c = 0; c < D; ++c) h[c] = fmaxf(h[c], 0.f);
    float f[D];
    load_bias<D>(f, b2);
    dense_acc<D, D>(w2, h, f);
    store_row64(F + r * D, f);
  }
}

// q[n] = max over the V views of F[n][v]   (transformer_network.py:286: no mask)
__global__ void __launch_bounds__(256) k_gnt_qinit(size_t n_elems, int V, const float* __restrict__ F, float* __restrict__ q) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / D;
    const int c = (int)(i - n * D);
    float m = -3.4e38f;
    for (int v = 0; v < V; ++v) m = fmaxf(m, F[(n * V + v) * D + c]);
    q[i] = m;
  }
}

// ---------------------------------------------------------------------------------------------------
// view transformer attention block (Transformer2D first half + Attention2D, :55-113), one thread per sample:
//   x = LN(q);  qq = q_fc(x);  per view: k = k_fc(F_v), v = v_fc(k), pos = pos_fc(ray_diff_v),
//   a = attn_fc(k - qq + pos) (masked_fill -1e9), softmax over views PER CHANNEL, out = sum_v (v + pos) a;
//   q <- out_fc(out) + q.   The softmax over views runs online (running max / sum / weighted accumulator).
// ---------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128) k_gnt_view_attn(int N, int V, const float* __restrict__ F,
                                                        const float* __restrict__ ray_diff, const float* __restrict__ mask,
                                                        const float* __restrict__ lp, const float* q_in, float* q,
                                                        float* __restrict__ vp_out = nullptr, float* __restrict__ a8_out = nullptr) {
  // vp_out / a8_out (nfb_gnt_bwd's checkpointing forward): per (sample, view) row v + pos [64] and ReLU(attn_fc.0(k - qq + pos)) [8]
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, nt = blockDim.x;
  load_vec_padded(sm + VS_LN_W, lp + L_V_LN1_W, D, D, t, nt);
  load_vec_padded(sm + VS_LN_B, lp + L_V_LN1_B, D, D, t, nt);
  load_wt_transposed(sm + VS_Q, lp + L_V_Q, D, D, D, t, nt);
  load_wt_transposed(sm + VS_K, lp + L_V_K, D, D, D, t, nt);
  load_wt_transposed(sm + VS_V, lp + L_V_V, D, D, D, t, nt);
  load_wt_transposed(sm + VS_O, lp + L_V_O_W, D, D, D, t, nt);
  load_vec_padded(sm + VS_O_B, lp + L_V_O_B, D, D, t, nt);
  load_wt_transposed(sm + VS_P0, lp + L_V_POS0_W, 8, 4, 8, t, nt);
  load_vec_padded(sm + VS_P0_B, lp + L_V_POS0_B, 8, 8, t, nt);
  load_wt_transposed(sm + VS_P2, lp + L_V_POS2_W, D, 8, D, t, nt);
  load_vec_padded(sm + VS_P2_B, lp + L_V_POS2_B, D, D, t, nt);
  load_wt_transposed(sm + VS_A0, lp + L_V_AT0_W, 8, D, 8, t, nt);
  load_vec_padded(sm + VS_A0_B, lp + L_V_AT0_B, 8, 8, t, nt);
  load_wt_transposed(sm + VS_A2, lp + L_V_AT2_W, D, 8, D, t, nt);
  load_vec_padded(sm + VS_A2_B, lp + L_V_AT2_B, D, D, t, nt);
  __syncthreads();

  for (int n = blockIdx.x * blockDim.x + t; n < N; n += gridDim.x * blockDim.x) {
    float q0[D];
    load_row64(q_in + (size_t)n * D, q0);
    float qq[D];
    {
      float x[D];
      layer_norm64(q0, sm + VS_LN_W, sm + VS_LN_B, LN_EPS_T, x);
#pragma unroll
      for (int c = 0; c < D; ++c) qq[c] = 0.f;
      dense_acc<D, D>(sm + VS_Q, x, qq);
    }
    float m[D], l[D], acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) { m[c] = -3.4e38f; l[c] = 0.f; acc[c] = 0.f; }
    for (int v = 0; v < V; ++v) {
      const size_t row = (size_t)n * V + v;
      float k[D], vv[D], pos[D];
      {
        float f[D];
        load_row64(F + row * D, f);
#pragma unroll
        for (int c = 0; c < D; ++c) k[c] = 0.f;
        dense_acc<D, D>(sm + VS_K, f, k);
      }
#pragma unroll
      for (int c = 0; c < D; ++c) vv[c] = 0.f;
      dense_acc<D, D>(sm + VS_V, k, vv);              // v = v_fc(k): applied to the projected k (:76-77)
      {
        const float4 rd4 = __ldg(reinterpret_cast<const float4*>(ray_diff) + row);
        const float rd[4] = {rd4.x, rd4.y, rd4.z, rd4.w};
        float p8[8];
        load_bias<8>(p8, sm + VS_P0_B);
        dense_acc<4, 8>(sm + VS_P0, rd, p8);
#pragma unroll
        for (int j = 0; j < 8; ++j) p8[j] = fmaxf(p8[j], 0.f);
        load_bias<D>(pos, sm + VS_P2_B);
        dense_acc<8, D>(sm + VS_P2, p8, pos);
      }
      float a8[8];
      load_bias<8>(a8, sm + VS_A0_B);
#pragma unroll
      for (int c = 0; c < D; ++c) axpy_row<8>(a8, k[c] - qq[c] + pos[c], sm + VS_A0 + c * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) a8[j] = fmaxf(a8[j], 0.f);
      if (vp_out) {
        float vp[D];
#pragma unroll
        for (int c = 0; c < D; ++c) vp[c] = vv[c] + pos[c];
        store_row64(vp_out + row * D, vp);
        float4* o8 = reinterpret_cast<float4*>(a8_out + row * 8);
        o8[0] = make_float4(a8[0], a8[1], a8[2], a8[3]);
        o8[1] = make_float4(a8[4], a8[5], a8[6], a8[7]);
      }
      const bool valid = __ldg(mask + row) != 0.f;
      float a[D];
      load_bias<D>(a, sm + VS_A2_B);
      dense_acc<8, D>(sm + VS_A2, a8, a);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const float s = valid ? a[c] : -1e9f;
        const float mn = fmaxf(m[c], s);
        const float sc = __expf(m[c] - mn), e = __expf(s - mn);
        l[c] = fmaf(l[c], sc, e);
        acc[c] = fmaf(acc[c], sc, (vv[c] + pos[c]) * e);
        m[c] = mn;
      }
    }
    float o[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = acc[c] / l[c];
    load_bias<D>(o, sm + VS_O_B);
    dense_acc<D, D>(sm + VS_O, acc, o);
#pragma unroll
    for (int c = 0; c < D; ++c) o[c] += q0[c];
    store_row64(q + (size_t)n * D, o);
  }
}

// ---------------------------------------------------------------------------------------------------
// feed-forward block (second half of Transformer2D / Transformer): q <- fc2(ReLU(fc1(LN(q)))) + q
// lp points at the block's {ff_norm.w, ff_norm.b, fc1.w, fc1.b, fc2.w, fc2.b} (contiguous in the blob)
// ---------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256, 1) k_gnt_ffn(int N, const float* __restrict__ lp, const float* q_in, float* q) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, nt = blockDim.x;
  load_vec_padded(sm + FS_LN_W, lp, D, D, t, nt);
  load_vec_padded(sm + FS_LN_B, lp + D, D, D, t, nt);
  load_wt_transposed(sm + FS_W1, lp + 2 * D, DH, D, DH, t, nt);
  load_vec_padded(sm + FS_B1, lp + 2 * D + DH * D, DH, DH, t, nt);
  load_wt_transposed(sm + FS_W2, lp + 2 * D + DH * D + DH, D, DH, D, t, nt);
  load_vec_padded(sm + FS_B2, lp + 2 * D + DH * D + DH + D * DH, D, D, t, nt);
  __syncthreads();
  for (int n = blockIdx.x * blockDim.x + t; n < N; n += gridDim.x * blockDim.x) {
    float q0[D], x[D], y[D];
    load_row64(q_in + (size_t)n * D, q0);
    layer_norm64(q0, sm + FS_LN_W, sm + FS_LN_B, LN_EPS_T, x);
    load_bias<D>(y, sm + FS_B2);
#pragma unroll 1
    for (int j0 = 0; j0 < DH; j0 += 32) {
      float h[32];
      load_bias<32>(h, sm + FS_B1 + j0);
#pragma unroll
      for (int k = 0; k < D; ++k) axpy_row<32>(h, x[k], sm + FS_W1 + k * DH + j0);
#pragma unroll
      for (int j = 0; j < 32; ++j) axpy_row<D>(y, fmaxf(h[j], 0.f), sm + FS_W2 + (j0 + j) * D);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) y[c] += q0[c];
    store_row64(q + (size_t)n * D, y);
  }
}

// ---------------------------------------------------------------------------------------------------
// q_fc on even layers (:295-297): q <- fc2(ReLU(fc1([q, posenc(pts), posenc(ray_d / |ray_d|)])))
// Embedder order (:12-37): [x, sin(x f0), cos(x f0), sin(x f1), ...], f_k = 2^k, k = 0..9, 3 components each.
// ---------------------------------------------------------------------------------------------------


__global__ void __launch_bounds__(128) k_gnt_qfc(int N, int S, const float* __restrict__ pts, const float* __restrict__ ray_d,
                                                  const float* __restrict__ lp, const float* q_in, float* q) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, nt = blockDim.x;
  load_wt_transposed(sm + QS_W0, lp + L_Q0_W, D, QIN, D, t, nt);
  load_vec_padded(sm + QS_B0, lp + L_Q0_B, D, D, t, nt);
  load_wt_transposed(sm + QS_W2, lp + L_Q2_W, D, D, D, t, nt);
  load_vec_padded(sm + QS_B2, lp + L_Q2_B, D, D, t, nt);
  __syncthreads();
  for (int n = blockIdx.x * blockDim.x + t; n < N; n += gridDim.x * blockDim.x) {
    float h[D];
    load_bias<D>(h, sm + QS_B0);
    {
      float q0[D];
      load_row64(q_in + (size_t)n * D, q0);
      dense_acc<D, D>(sm + QS_W0, q0, h);
    }
    const float p3[3] = {__ldg(pts + (size_t)n * 3), __ldg(pts + (size_t)n * 3 + 1), __ldg(pts + (size_t)n * 3 + 2)};
    posenc_axpy(h, p3, sm + QS_W0 + D * D);
    const int r = n / S;
    const float dx = __ldg(ray_d + (size_t)r * 3), dy = __ldg(ray_d + (size_t)r * 3 + 1), dz = __ldg(ray_d + (size_t)r * 3 + 2);
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    const float d3[3] = {__fdiv_rn(dx, nrm), __fdiv_rn(dy, nrm), __fdiv_rn(dz, nrm)};
    posenc_axpy(h, d3, sm + QS_W0 + (D + PE) * D);
#pragma unroll
    for (int c = 0; c < D; ++c) h[c] = fmaxf(h[c], 0.f);
    float y[D];
    load_bias<D>(y, sm + QS_B2);
    dense_acc<D, D>(sm + QS_W2, h, y);
    store_row64(q + (size_t)n * D, y);
  }
}

// ---------------------------------------------------------------------------------------------------
// ray transformer attention block (Transformer first half + Attention "qk", :121-202): one CTA per ray, one thread per
// sample; K / V rows of the ray in shared memory; 4 heads x 16, scores / sqrt(16), softmax over the samples, no mask.
// q <- out_fc(attn) + q.   attn_out (optional): mean over heads of the probabilities of QUERY 0 (:200) -> [R][S].
// ---------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256, 1) k_gnt_ray_attn(int R, int S, int rpc, const float* __restrict__ lp, const float* q_in, float* q,
                                                          float* __restrict__ attn_out, int attn_stride) {
  extern __shared__ __align__(16) float sm[];
  const int nt = blockDim.x;
  const int rb = nt / rpc;               // threads per ray (S rounded up to a warp multiple)
  const int lr = threadIdx.x / rb;       // ray of the CTA this thread works on
  const int t = threadIdx.x - lr * rb;   // sample index
  float* sq0 = sm + RS_W_TOTAL + (size_t)lr * (RS_PER_RAY + 2 * S * D);   // this ray's query-0 block
  float* sk = sq0 + RS_PER_RAY;         // [S][64]
  float* sv = sk + (size_t)S * D;       // [S][64]
  load_vec_padded(sm + RS_LN_W, lp + L_R_LN1_W, D, D, threadIdx.x, nt);
  load_vec_padded(sm + RS_LN_B, lp + L_R_LN1_B, D, D, threadIdx.x, nt);
  load_wt_transposed(sm + RS_Q, lp + L_R_Q, D, D, D, threadIdx.x, nt);
  load_wt_transposed(sm + RS_K, lp + L_R_K, D, D, D, threadIdx.x, nt);
  load_wt_transposed(sm + RS_V, lp + L_R_V, D, D, D, threadIdx.x, nt);
  load_wt_transposed(sm + RS_O, lp + L_R_O_W, D, D, D, threadIdx.x, nt);
  load_vec_padded(sm + RS_O_B, lp + L_R_O_B, D, D, threadIdx.x, nt);
  __syncthreads();
  for (int r0 = blockIdx.x * rpc; r0 < R; r0 += gridDim.x * rpc) {
    const int r = r0 + lr;
    const bool act = (t < S) && (r < R);
    const size_t qoff = ((size_t)(r < R ? r : 0) * S + (t < S ? t : 0)) * D;
    float* qrow = q + qoff;
    float q0[D], qv[D];
    load_row64(q_in + qoff, q0);
    {
      float x[D], kk[D];
      layer_norm64(q0, sm + RS_LN_W, sm + RS_LN_B, LN_EPS_T, x);
#pragma unroll
      for (int c = 0; c < D; ++c) { qv[c] = 0.f; kk[c] = 0.f; }
      dense_acc<D, D>(sm + RS_Q, x, qv);
      dense_acc<D, D>(sm + RS_K, x, kk);
      if (act) store_row64(sk + (size_t)t * D, kk);
#pragma unroll
      for (int c = 0; c < D; ++c) kk[c] = 0.f;
      dense_acc<D, D>(sm + RS_V, x, kk);
      if (act) store_row64(sv + (size_t)t * D, kk);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) qv[c] *= 0.25f;          // 1 / sqrt(16)
    if (t == 0) store_row64(sq0 + RS_Q0, qv);
    __syncthreads();
    float o[D];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float mx = -3.4e38f;
      for (int j = 0; j < S; ++j) {
        const float* kj = sk + (size_t)j * D + 16 * h;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) s = fmaf(qv[16 * h + c], kj[c], s);
        mx = fmaxf(mx, s);
      }
      float l = 0.f, a16[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) a16[c] = 0.f;
      for (int j = 0; j < S; ++j) {
        const float* kj = sk + (size_t)j * D + 16 * h;
        const float* vj = sv + (size_t)j * D + 16 * h;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) s = fmaf(qv[16 * h + c], kj[c], s);
        const float p = __expf(s - mx);
        l += p;
#pragma unroll
        for (int c = 0; c < 16; ++c) a16[c] = fmaf(p, vj[c], a16[c]);
      }
      const float il = 1.f / l;
#pragma unroll
      for (int c = 0; c < 16; ++c) o[16 * h + c] = a16[c] * il;
      if (t == 0) { sq0[RS_ST + h] = mx; sq0[RS_ST + 4 + h] = il; }
    }
    float y[D];
    load_bias<D>(y, sm + RS_O_B);
    dense_acc<D, D>(sm + RS_O, o, y);
#pragma unroll
    for (int c = 0; c < D; ++c) y[c] += q0[c];
    if (act) store_row64(qrow, y);
    if (attn_out) {
      __syncthreads();                                     // query 0's statistics are published
      if (act) {
        float pm = 0.f;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float* kj = sk + (size_t)t * D + 16 * h;
          float s = 0.f;
#pragma unroll
          for (int c = 0; c < 16; ++c) s = fmaf(sq0[RS_Q0 + 16 * h + c], kj[c], s);
          pm += __expf(s - sq0[RS_ST + h]) * sq0[RS_ST + 4 + h];
        }
        attn_out[(size_t)r * attn_stride + t] = 0.25f * pm;
      }
    }
    __syncthreads();                                       // sk / sv / q0 are rewritten by the next ray
  }
}

// ---------------------------------------------------------------------------------------------------
// head (:303-305): h = LayerNorm(q) (eps 1e-5), rgb = rgb_fc(mean over samples of h).  One CTA per ray.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gnt_head(int R, int S, const float* __restrict__ tp, const float* __restrict__ q,
                                                   float* __restrict__ out, int out_stride) {
  extern __shared__ __align__(16) float sm[];
  float* sh = sm;                        // [S][65]
  float* smean = sm + (size_t)S * 65;    // [64]
  const int t = threadIdx.x;
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    if (t < S) {
      float q0[D], h[D];
      load_row64(q + ((size_t)r * S + t) * D, q0);
      float w[D], b[D];
#pragma unroll
      for (int c = 0; c < D; ++c) { w[c] = __ldg(tp + T_LN_W + c); b[c] = __ldg(tp + T_LN_B + c); }
      layer_norm64(q0, w, b, LN_EPS_HEAD, h);
#pragma unroll
      for (int c = 0; c < D; ++c) sh[t * 65 + c] = h[c];
    }
    __syncthreads();
    if (t < D) {
      float s = 0.f;
      for (int j = 0; j < S; ++j) s += sh[j * 65 + t];
      smean[t] = s / (float)S;
    }
    __syncthreads();
    if (t < 3) {
      float a = __ldg(tp + T_RGB_B + t);
      for (int c = 0; c < D; ++c) a = fmaf(smean[c], __ldg(tp + T_RGB_W + t * D + c), a);
      out[(size_t)r * out_stride + t] = a;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// tensor-core form: the 64 x 64 projections run in k_gnt_lin_tc (nfb_gnt_tc.cuh); what stays on the CUDA cores is
// the part of the two attentions that is not a dense contraction.
// ---------------------------------------------------------------------------------------------------
// Input per (sample, view) row from k_gnt_lin_tc<LIN_KV>: a8 = ReLU(attn_fc.0(k - qq + pos)) [8] and vp = v + pos [64].
// Two adjacent lanes share a sample, each owns 32 of the 64 channels: a = attn_fc.2(a8) (masked), softmax over the
// views per channel (online), x = sum_v vp a.
__global__ void __launch_bounds__(256) k_gnt_view_core(int N, int V, const float* __restrict__ A8, const float* __restrict__ VP,
                                                        const float* __restrict__ mask, const float* __restrict__ lp,
                                                        float* __restrict__ xout) {
  // Eight adjacent lanes share a sample, each owns 8 of the 64 channels: a warp-wide load of the v + pos rows is 4 whole
  // 256-byte rows (with two lanes per sample it touched 16 rows for the same bytes), and the per-lane state is 24 registers.
  constexpr int LC = 8;
  __shared__ __align__(16) float sm[8 * D + D];          // attn_fc.2 transposed [8][64], bias [64]
  const int t = threadIdx.x, nt = blockDim.x;
  load_wt_transposed(sm, lp + L_V_AT2_W, D, 8, D, t, nt);
  load_vec_padded(sm + 8 * D, lp + L_V_AT2_B, D, D, t, nt);
  __syncthreads();
  const int c0 = (t & 7) * LC;
  for (int n = blockIdx.x * (blockDim.x / 8) + (t >> 3); n < N; n += gridDim.x * (blockDim.x / 8)) {
    float m[LC], l[LC], acc[LC];
#pragma unroll
    for (int c = 0; c < LC; ++c) { m[c] = -3.4e38f; l[c] = 0.f; acc[c] = 0.f; }
    for (int v = 0; v < V; ++v) {
      const size_t row = (size_t)n * V + v;
      const float4 h0 = __ldg(reinterpret_cast<const float4*>(A8 + row * 8)), h1 = __ldg(reinterpret_cast<const float4*>(A8 + row * 8) + 1);
      const float a8[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
      const bool valid = __ldg(mask + row) != 0.f;
      float a[LC];
      load_bias<LC>(a, sm + 8 * D + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) axpy_row<LC>(a, a8[j], sm + j * D + c0);
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(VP + row * D + c0)), v1 = __ldg(reinterpret_cast<const float4*>(VP + row * D + c0) + 1);
      const float vv[LC] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int c = 0; c < LC; ++c) {
        const float s_ = valid ? a[c] : -1e9f;
        const float mn = fmaxf(m[c], s_);
        const float sc = __expf(m[c] - mn), e = __expf(s_ - mn);
        l[c] = fmaf(l[c], sc, e);
        acc[c] = fmaf(acc[c], sc, vv[c] * e);
        m[c] = mn;
      }
    }
    float4* o = reinterpret_cast<float4*>(xout + (size_t)n * D + c0);
    o[0] = make_float4(acc[0] / l[0], acc[1] / l[1], acc[2] / l[2], acc[3] / l[3]);
    o[1] = make_float4(acc[4] / l[4], acc[5] / l[5], acc[6] / l[6], acc[7] / l[7]);
  }
}

// ---------------------------------------------------------------------------------------------------
// fp32 view attention in row-parallel pieces (the checkpointing forward of nfb_gnt_bwd): one thread per sample for the query side
// (k_gnt_pre, k_gnt_outfc), one thread per (sample, view) row for the key side (k_gnt_view_row_fwd: V x the parallelism of the fused
// k_gnt_view_attn and a third of its live state), k_gnt_view_core for the softmax over views.  Same arithmetic per value.
// ---------------------------------------------------------------------------------------------------
// qq[n] = q_fc(LN(q[n]))
__global__ void __launch_bounds__(128) k_gnt_pre(int N, const float* __restrict__ lp, const float* __restrict__ q_in, float* __restrict__ qq_out) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, nt = blockDim.x;
  load_vec_padded(sm, lp + L_V_LN1_W, D, D, t, nt);
  load_vec_padded(sm + D, lp + L_V_LN1_B, D, D, t, nt);
  load_wt_transposed(sm + 2 * D, lp + L_V_Q, D, D, D, t, nt);
  __syncthreads();
  for (int n = blockIdx.x * blockDim.x + t; n < N; n += gridDim.x * blockDim.x) {
    float x[D], qq[D];
    {
      float q0[D];
      load_row64(q_in + (size_t)n * D, q0);
      layer_norm64(q0, sm, sm + D, LN_EPS_T, x);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) qq[c] = 0.f;
    dense_acc<D, D>(sm + 2 * D, x, qq);
    store_row64(qq_out + (size_t)n * D, qq);
  }
}
// q_out[n] = out_fc(x[n]) + bias + q_in[n]
__global__ void __launch_bounds__(128) k_gnt_outfc(int N, const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ x_in,
                                                    const float* q_in, float* q_out) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, nt = blockDim.x;
  load_wt_transposed(sm, w, D, D, D, t, nt);
  load_vec_padded(sm + D * D, b, D, D, t, nt);
  __syncthreads();
  for (int n = blockIdx.x * blockDim.x + t; n < N; n += gridDim.x * blockDim.x) {
    float x[D], o[D];
    load_row64(x_in + (size_t)n * D, x);
    load_bias<D>(o, sm + D * D);
    dense_acc<D, D>(sm, x, o);
    float q0[D];
    load_row64(q_in + (size_t)n * D, q0);
#pragma unroll
    for (int c = 0; c < D; ++c) o[c] += q0[c];
    store_row64(q_out + (size_t)n * D, o);
  }
}
// per row: k = k_fc(F), v = v_fc(k), pos = pos_fc(ray_diff)  ->  VP = v + pos [64],  A8 = ReLU(attn_fc.0(k - qq + pos)) [8]
__global__ void __launch_bounds__(128) k_gnt_view_row_fwd(size_t rows, int V, const float* __restrict__ F, const float* __restrict__ QQ,
                                                           const float* __restrict__ ray_diff, const float* __restrict__ lp,
                                                           float* __restrict__ VP, float* __restrict__ A8) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, nt = blockDim.x;
  load_wt_transposed(sm + VS_K, lp + L_V_K, D, D, D, t, nt);
  load_wt_transposed(sm + VS_V, lp + L_V_V, D, D, D, t, nt);
  load_wt_transposed(sm + VS_P0, lp + L_V_POS0_W, 8, 4, 8, t, nt);
  load_vec_padded(sm + VS_P0_B, lp + L_V_POS0_B, 8, 8, t, nt);
  load_wt_transposed(sm + VS_P2, lp + L_V_POS2_W, D, 8, D, t, nt);
  load_vec_padded(sm + VS_P2_B, lp + L_V_POS2_B, D, D, t, nt);
  load_wt_transposed(sm + VS_A0, lp + L_V_AT0_W, 8, D, 8, t, nt);
  load_vec_padded(sm + VS_A0_B, lp + L_V_AT0_B, 8, 8, t, nt);
  __syncthreads();
  for (size_t row = (size_t)blockIdx.x * blockDim.x + t; row < rows; row += (size_t)gridDim.x * blockDim.x) {
    float k[D], vv[D], pos[D];
    {
      float f[D];
      load_row64(F + row * D, f);
#pragma unroll
      for (int c = 0; c < D; ++c) k[c] = 0.f;
      dense_acc<D, D>(sm + VS_K, f, k);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) vv[c] = 0.f;
    dense_acc<D, D>(sm + VS_V, k, vv);
    {
      const float4 rd4 = __ldg(reinterpret_cast<const float4*>(ray_diff) + row);
      const float rd[4] = {rd4.x, rd4.y, rd4.z, rd4.w};
      float p8[8];
      load_bias<8>(p8, sm + VS_P0_B);
      dense_acc<4, 8>(sm + VS_P0, rd, p8);
#pragma unroll
      for (int j = 0; j < 8; ++j) p8[j] = fmaxf(p8[j], 0.f);
      load_bias<D>(pos, sm + VS_P2_B);
      dense_acc<8, D>(sm + VS_P2, p8, pos);
    }
    float a8[8];
    load_bias<8>(a8, sm + VS_A0_B);
    {
      float qq[D];
      load_row64(QQ + (row / V) * D, qq);
#pragma unroll
      for (int c = 0; c < D; ++c) axpy_row<8>(a8, k[c] - qq[c] + pos[c], sm + VS_A0 + c * 8);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) vv[c] += pos[c];
    store_row64(VP + row * D, vv);
    float4* o8 = reinterpret_cast<float4*>(A8 + row * 8);
    o8[0] = make_float4(fmaxf(a8[0], 0.f), fmaxf(a8[1], 0.f), fmaxf(a8[2], 0.f), fmaxf(a8[3], 0.f));
    o8[1] = make_float4(fmaxf(a8[4], 0.f), fmaxf(a8[5], 0.f), fmaxf(a8[6], 0.f), fmaxf(a8[7], 0.f));
  }
}

// ray core: scaled-dot-product attention of the ray's samples given the projected Q, K, V [N][64]; o -> [N][64]
// (same CTA / thread mapping as k_gnt_ray_attn; the projections and out_fc run in k_gnt_lin_tc)
__global__ void __launch_bounds__(256, 1) k_gnt_ray_core(int R, int S, int rpc, const float* __restrict__ Q, const float* __restrict__ K,
                                                          const float* __restrict__ Vp, float* __restrict__ O,
                                                          float* __restrict__ attn_out, int attn_stride) {
  extern __shared__ __align__(16) float sm[];
  const int nt = blockDim.x;
  const int rb = nt / rpc;
  const int lr = threadIdx.x / rb;
  const int t = threadIdx.x - lr * rb;
  float* sq0 = sm + (size_t)lr * (RS_PER_RAY + 2 * S * D);
  float* sk = sq0 + RS_PER_RAY;
  float* sv = sk + (size_t)S * D;
  for (int r0 = blockIdx.x * rpc; r0 < R; r0 += gridDim.x * rpc) {
    const int r = r0 + lr;
    const bool act = (t < S) && (r < R);
    const size_t n = (size_t)(r < R ? r : 0) * S + (t < S ? t : 0);
    float qv[D];
    load_row64(Q + n * D, qv);
    if (r < R) {
      // the ray's K / V rows are contiguous in memory and laid out like sk / sv: coalesced block copy by the ray's threads
      const float4* k4 = reinterpret_cast<const float4*>(K + (size_t)r * S * D);
      const float4* v4 = reinterpret_cast<const float4*>(Vp + (size_t)r * S * D);
      float4* sk4 = reinterpret_cast<float4*>(sk);
      float4* sv4 = reinterpret_cast<float4*>(sv);
      for (int i = t; i < S * (D / 4); i += rb) { sk4[i] = __ldg(k4 + i); sv4[i] = __ldg(v4 + i); }
    }
#pragma unroll
    for (int c = 0; c < D; ++c) qv[c] *= 0.25f;          // 1 / sqrt(16)
    if (t == 0) store_row64(sq0 + RS_Q0, qv);
    __syncthreads();
    float o[D];
    const int Sr = (r < R) ? S : 0;
    // one pass over the keys per head: online softmax (running max, rescale on the rare increase) and packed fp32x2
    // arithmetic (FFMA2 issues at twice the scalar FFMA rate on sm_100: tests/probes/probe_ffma2.cu)
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float2 q2[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) q2[c] = make_float2(qv[16 * h + 2 * c], qv[16 * h + 2 * c + 1]);
      float mx = -3.4e38f, l = 0.f;
      float2 a2[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) a2[c] = make_float2(0.f, 0.f);
      for (int j = 0; j < Sr; ++j) {
        const float4* kj = reinterpret_cast<const float4*>(sk + (size_t)j * D + 16 * h);
        const float4* vj = reinterpret_cast<const float4*>(sv + (size_t)j * D + 16 * h);
        float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 k4 = kj[c];
          s2 = __ffma2_rn(q2[2 * c], make_float2(k4.x, k4.y), s2);
          s2 = __ffma2_rn(q2[2 * c + 1], make_float2(k4.z, k4.w), s2);
        }
        const float sc = s2.x + s2.y;
        if (sc > mx) {                                     // new maximum: rescale what has been accumulated
          const float r = __expf(mx - sc);
          l *= r;
#pragma unroll
          for (int c = 0; c < 8; ++c) a2[c] = __fmul2_rn(a2[c], make_float2(r, r));
          mx = sc;
        }
        const float p = __expf(sc - mx);
        l += p;
        const float2 p2 = make_float2(p, p);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 v4 = vj[c];
          a2[2 * c] = __ffma2_rn(p2, make_float2(v4.x, v4.y), a2[2 * c]);
          a2[2 * c + 1] = __ffma2_rn(p2, make_float2(v4.z, v4.w), a2[2 * c + 1]);
        }
      }
      const float il = 1.f / l;
#pragma unroll
      for (int c = 0; c < 8; ++c) { o[16 * h + 2 * c] = a2[c].x * il; o[16 * h + 2 * c + 1] = a2[c].y * il; }
      if (t == 0) { sq0[RS_ST + h] = mx; sq0[RS_ST + 4 + h] = il; }
    }
    __syncthreads();                                       // every query of the ray is done with sv; query 0's statistics are published
    if (act) store_row64(sv + (size_t)t * D, o);           // o rows leave through sv: coalesced block copy below
    if (attn_out && act) {
      float pm = 0.f;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float* kj = sk + (size_t)t * D + 16 * h;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) s = fmaf(sq0[RS_Q0 + 16 * h + c], kj[c], s);
        pm += __expf(s - sq0[RS_ST + h]) * sq0[RS_ST + 4 + h];
      }
      attn_out[(size_t)r * attn_stride + t] = 0.25f * pm;
    }
    __syncthreads();
    if (r < R) {
      float4* o4 = reinterpret_cast<float4*>(O + (size_t)r * S * D);
      const float4* sv4 = reinterpret_cast<const float4*>(sv);
      for (int i = t; i < S * (D / 4); i += rb) o4[i] = sv4[i];
    }
    __syncthreads();
  }
}

template <typename K>
int set_smem(K kern, size_t bytes, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
  return NFB_OK;
}

}  // namespace

extern "C" int nfb_gnt_param_floats(int depth) { return depth < 1 ? 0 : G_HEAD + depth * L_SIZE + T_SIZE; }

extern "C" int nfb_gnt_param_offset(int depth, const char* name) {
  // offsets of the header / tail tensors and of the per-layer block, for the host-side packer and its test
  if (!name) return -1;
  struct E { const char* n; int off; };
  static const E head[] = {{"rgbfeat_fc.0.weight", G_RF0_W}, {"rgbfeat_fc.0.bias", G_RF0_B},
                           {"rgbfeat_fc.2.weight", G_RF2_W}, {"rgbfeat_fc.2.bias", G_RF2_B}, {"layer0", G_HEAD}};
  for (const E& e : head)
    if (strcmp(e.n, name) == 0) return e.off;
  static const E lay[] = {
      {"view.attn_norm.weight", L_V_LN1_W}, {"view.attn_norm.bias", L_V_LN1_B}, {"view.attn.q_fc.weight", L_V_Q},
      {"view.attn.k_fc.weight", L_V_K}, {"view.attn.v_fc.weight", L_V_V}, {"view.attn.pos_fc.0.weight", L_V_POS0_W},
      {"view.attn.pos_fc.0.bias", L_V_POS0_B}, {"view.attn.pos_fc.2.weight", L_V_POS2_W}, {"view.attn.pos_fc.2.bias", L_V_POS2_B},
      {"view.attn.attn_fc.0.weight", L_V_AT0_W}, {"view.attn.attn_fc.0.bias", L_V_AT0_B},
      {"view.attn.attn_fc.2.weight", L_V_AT2_W}, {"view.attn.attn_fc.2.bias", L_V_AT2_B},
      {"view.attn.out_fc.weight", L_V_O_W}, {"view.attn.out_fc.bias", L_V_O_B}, {"view.ff_norm.weight", L_V_LN2_W},
      {"view.ff_norm.bias", L_V_LN2_B}, {"view.ff.fc1.weight", L_V_FF1_W}, {"view.ff.fc1.bias", L_V_FF1_B},
      {"view.ff.fc2.weight", L_V_FF2_W}, {"view.ff.fc2.bias", L_V_FF2_B}, {"q_fc.0.weight", L_Q0_W}, {"q_fc.0.bias", L_Q0_B},
      {"q_fc.2.weight", L_Q2_W}, {"q_fc.2.bias", L_Q2_B}, {"ray.attn_norm.weight", L_R_LN1_W}, {"ray.attn_norm.bias", L_R_LN1_B},
      {"ray.attn.q_fc.weight", L_R_Q}, {"ray.attn.k_fc.weight", L_R_K}, {"ray.attn.v_fc.weight", L_R_V},
      {"ray.attn.out_fc.weight", L_R_O_W}, {"ray.attn.out_fc.bias", L_R_O_B}, {"ray.ff_norm.weight", L_R_LN2_W},
      {"ray.ff_norm.bias", L_R_LN2_B}, {"ray.ff.fc1.weight", L_R_FF1_W}, {"ray.ff.fc1.bias", L_R_FF1_B},
      {"ray.ff.fc2.weight", L_R_FF2_W}, {"ray.ff.fc2.bias", L_R_FF2_B}, {"layer_size", L_SIZE}};
  for (const E& e : lay)
    if (strcmp(e.n, name) == 0) return e.off;
  const int tail = G_HEAD + depth * L_SIZE;
  static const E tl[] = {{"norm.weight", T_LN_W}, {"norm.bias", T_LN_B}, {"rgb_fc.weight", T_RGB_W}, {"rgb_fc.bias", T_RGB_B}};
  for (const E& e : tl)
    if (strcmp(e.n, name) == 0) return tail + e.off;
  return -1;
}

extern "C" size_t nfb_gnt_workspace_bytes(int R, int S, int V) {
  // F[rows][64] + q[N][64]  |  tensor-core form in addition: k, v [rows][64] and five per-sample [N][64] buffers
  if (R <= 0 || S < 1 || V < 1) return 0;
  return ((size_t)R * S * V * D * 3 + (size_t)R * S * D * 6) * sizeof(float);
}

template <int NPASS>
static int gnt_layer_tc(int i, int R, int S, int V, const float* ray_diff, const float* mask, const float* pts, const float* ray_d,
                        const float* lp, float* F, float* q, float* ws, float* attn_out, int out_stride, int rpc, int ray_grid,
                        int ray_block, int sms, cudaStream_t st) {
  using namespace gnttc;
  const int N = R * S;
  const size_t rows = (size_t)N * V;
  float* Kv = ws;                       // [rows][64]
  float* Vv = Kv + rows * D;            // [rows][64]
  float* b0 = Vv + rows * D;            // five [N][64] per-sample buffers
  float* b1 = b0 + (size_t)N * D;
  float* b2 = b1 + (size_t)N * D;
  float* b3 = b2 + (size_t)N * D;
  int rc;
  // ---- view transformer: qq = q_fc(LN(q)); k = k_fc(F), v = v_fc(k); core; q = out_fc(x) + b + q; FFN
  LinArgs a{};
  a.M = N; a.x = q; a.y0 = b0; a.w[0] = lp + L_V_Q; a.ln_w = lp + L_V_LN1_W; a.ln_b = lp + L_V_LN1_B;
  if ((rc = launch_lin<NPASS, LIN_PRE>(a, st, "k_gnt_lin_tc<pre>"))) return rc;
  a = LinArgs{};
  a.M = (long long)rows; a.x = F; a.y0 = Vv; a.y1 = Kv; a.w[0] = lp + L_V_K; a.w[1] = lp + L_V_V;      // y0 = v + pos, y1 = a8
  a.qq = b0; a.ray_diff = ray_diff; a.V = V;
  a.p0_w = lp + L_V_POS0_W; a.p0_b = lp + L_V_POS0_B; a.p2_w = lp + L_V_POS2_W; a.p2_b = lp + L_V_POS2_B;
  a.a0_w = lp + L_V_AT0_W; a.a0_b = lp + L_V_AT0_B;
  if ((rc = launch_lin<NPASS, LIN_KV>(a, st, "k_gnt_lin_tc<kv>"))) return rc;
  {
    int g = (N + 31) / 32;                      // 8 threads per sample, 256 per CTA
    if (g > sms * 8) g = sms * 8;
    k_gnt_view_core<<<g, 256, 0, st>>>(N, V, Kv, Vv, mask, lp, b1);
    NFB_CHECK_LAUNCH("k_gnt_view_core");
  }
  a = LinArgs{};
  a.M = N; a.x = b1; a.res = q; a.y0 = q; a.w[0] = lp + L_V_O_W; a.b0 = lp + L_V_O_B;
  if ((rc = launch_lin<NPASS, LIN_POST>(a, st, "k_gnt_lin_tc<post>"))) return rc;
  a = LinArgs{};
  a.M = N; a.x = q; a.y0 = q; a.w[0] = lp + L_V_FF1_W; a.w[1] = lp + L_V_FF2_W; a.b0 = lp + L_V_FF1_B; a.b1 = lp + L_V_FF2_B;
  a.ln_w = lp + L_V_LN2_W; a.ln_b = lp + L_V_LN2_B;
  if ((rc = launch_lin<NPASS, LIN_FFN>(a, st, "k_gnt_lin_tc<ffn>"))) return rc;
  if ((i & 1) == 0) {
    a = LinArgs{};
    a.M = N; a.x = q; a.y0 = q; a.w[0] = lp + L_Q0_W; a.w[1] = lp + L_Q2_W; a.b0 = lp + L_Q0_B; a.b1 = lp + L_Q2_B;
    a.pts = pts; a.ray_d = ray_d; a.S = S;
    if ((rc = launch_lin<NPASS, LIN_QFC>(a, st, "k_gnt_lin_tc<qfc>"))) return rc;
  }
  // ---- ray transformer: Q, K, V = projections of LN(q); core; q = out_fc(o) + b + q; FFN
  a = LinArgs{};
  a.M = N; a.x = q; a.y0 = b0; a.y1 = b1; a.y2 = b2; a.w[0] = lp + L_R_Q; a.w[1] = lp + L_R_K; a.w[2] = lp + L_R_V;
  a.ln_w = lp + L_R_LN1_W; a.ln_b = lp + L_R_LN1_B;
  if ((rc = launch_lin<NPASS, LIN_QKV>(a, st, "k_gnt_lin_tc<qkv>"))) return rc;
  k_gnt_ray_core<<<ray_grid, ray_block * rpc, (size_t)rpc * (2 * S * D + RS_PER_RAY) * sizeof(float), st>>>(
      R, S, rpc, b0, b1, b2, b3, attn_out, out_stride);
  NFB_CHECK_LAUNCH("k_gnt_ray_core");
  a = LinArgs{};
  a.M = N; a.x = b3; a.res = q; a.y0 = q; a.w[0] = lp + L_R_O_W; a.b0 = lp + L_R_O_B;
  if ((rc = launch_lin<NPASS, LIN_POST>(a, st, "k_gnt_lin_tc<post>"))) return rc;
  a = LinArgs{};
  a.M = N; a.x = q; a.y0 = q; a.w[0] = lp + L_R_FF1_W; a.w[1] = lp + L_R_FF2_W; a.b0 = lp + L_R_FF1_B; a.b1 = lp + L_R_FF2_B;
  a.ln_w = lp + L_R_LN2_W; a.ln_b = lp + L_R_LN2_B;
  return launch_lin<NPASS, LIN_FFN>(a, st, "k_gnt_lin_tc<ffn>");
}

extern "C" int nfb_gnt_fwd(int R, int S, int V, int depth, int ret_alpha, const float* rgb_feat, const float* ray_diff,
                           const float* mask, const float* pts, const float* ray_d, const float* params, float* out,
                           void* workspace, size_t workspace_bytes, int precision, void* stream) {
  NFB_REQUIRE(precision >= NFB_PREC_FP32 && precision <= NFB_PREC_BF16, NFB_EINVAL, "nfb_gnt_fwd: bad precision %d", precision);
  NFB_REQUIRE(R >= 0 && S >= 1 && V >= 1 && depth >= 1, NFB_EINVAL, "nfb_gnt_fwd: bad arguments (R=%d S=%d V=%d depth=%d)", R, S, V, depth);
  NFB_REQUIRE(S <= NFB_MAX_SAMPLES, NFB_EUNSUPPORTED, "nfb_gnt_fwd: S=%d > %d samples per ray", S, NFB_MAX_SAMPLES);
  NFB_REQUIRE(V <= NFB_MAX_VIEWS, NFB_EUNSUPPORTED, "nfb_gnt_fwd: V=%d > %d views", V, NFB_MAX_VIEWS);
  if (R == 0) return NFB_OK;
  NFB_REQUIRE(rgb_feat && ray_diff && mask && pts && ray_d && params && out && workspace, NFB_EINVAL, "nfb_gnt_fwd: NULL buffer");
  NFB_REQUIRE(workspace_bytes >= nfb_gnt_workspace_bytes(R, S, V), NFB_EINVAL, "nfb_gnt_fwd: workspace too small (%zu < %zu bytes)",
              workspace_bytes, nfb_gnt_workspace_bytes(R, S, V));
  NFB_REQUIRE(((uintptr_t)ray_diff % 16) == 0 && ((uintptr_t)workspace % 16) == 0 && ((uintptr_t)params % 16) == 0, NFB_EINVAL,
              "nfb_gnt_fwd: ray_diff / workspace / params must be 16-byte aligned");
  NFB_REQUIRE((long long)R * S * V < (1ll << 31), NFB_EUNSUPPORTED, "nfb_gnt_fwd: more than 2^31 rows per call; chunk the rays");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = R * S;
  const size_t rows = (size_t)N * V;
  float* F = reinterpret_cast<float*>(workspace);
  float* q = F + rows * D;
  const int sms = nfb_num_sms();
  int rc;

  const size_t sm_embed = (size_t)(35 * D + D + D * D + D) * sizeof(float);
  if ((rc = set_smem(k_gnt_embed, sm_embed, "k_gnt_embed"))) return rc;
  {
    if (precision != NFB_PREC_FP32) {
      // rgbfeat_fc as two tensor-core tiles (35 -> 64 zero-padded to K = 64, ReLU, 64 -> 64)
      gnttc::LinArgs a{};
      a.M = (long long)rows; a.x = rgb_feat; a.y0 = F; a.w[0] = params + G_RF0_W; a.w[1] = params + G_RF2_W;
      a.b0 = params + G_RF0_B; a.b1 = params + G_RF2_B;
      rc = precision == NFB_PREC_BF16 ? gnttc::launch_lin<1, gnttc::LIN_EMBED>(a, st, "k_gnt_lin_tc<embed>")
                                      : gnttc::launch_lin<3, gnttc::LIN_EMBED>(a, st, "k_gnt_lin_tc<embed>");
      if (rc) return rc;
    } else {
      size_t g = (rows + 127) / 128;
      if (g > (size_t)sms * 8) g = (size_t)sms * 8;
      k_gnt_embed<<<(int)g, 128, sm_embed, st>>>(rows, rgb_feat, params, F);
      NFB_CHECK_LAUNCH("k_gnt_embed");
    }
    size_t g2 = ((size_t)N * D + 255) / 256;
    if (g2 > (size_t)sms * 16) g2 = (size_t)sms * 16;
    k_gnt_qinit<<<(int)g2, 256, 0, st>>>((size_t)N * D, V, F, q);
    NFB_CHECK_LAUNCH("k_gnt_qinit");
  }
  const size_t sm_view = (size_t)VS_TOTAL * sizeof(float), sm_ffn = (size_t)FS_TOTAL * sizeof(float),
               sm_qfc = (size_t)QS_TOTAL * sizeof(float), sm_ray = (size_t)(RS_W_TOTAL + (size_t)(256 / (((S + 31) / 32) * 32) > 0 ? 256 / (((S + 31) / 32) * 32) : 1) * (2 * S * D + RS_PER_RAY)) * sizeof(float),
               sm_head = (size_t)(S * 65 + D) * sizeof(float);
  if ((rc = set_smem(k_gnt_view_attn, sm_view, "k_gnt_view_attn"))) return rc;
  if ((rc = set_smem(k_gnt_ffn, sm_ffn, "k_gnt_ffn"))) return rc;
  if ((rc = set_smem(k_gnt_qfc, sm_qfc, "k_gnt_qfc"))) return rc;
  if ((rc = set_smem(k_gnt_ray_attn, sm_ray, "k_gnt_ray_attn"))) return rc;
  if ((rc = set_smem(k_gnt_head, sm_head, "k_gnt_head"))) return rc;
  if ((rc = set_smem(k_gnt_ray_core, sm_ray, "k_gnt_ray_core"))) return rc;
  auto grid_for = [&](int per_sm) {
    int g = (N + 127) / 128;
    if (g > sms * per_sm) g = sms * per_sm;
    return g < 1 ? 1 : g;
  };
  const int ffn_grid = (N + 255) / 256 < sms ? (N + 255) / 256 : sms;
  const int ray_block = ((S + 31) / 32) * 32;
  // rays per CTA of the ray-attention kernel: up to 256 threads, K/V of every ray of the CTA in shared memory
  int rpc = 256 / ray_block;
  const int rpc_smem = (int)((200 * 1024 - (size_t)RS_W_TOTAL * sizeof(float)) / ((size_t)(2 * S * D + RS_PER_RAY) * sizeof(float)));
  if (rpc > rpc_smem) rpc = rpc_smem;
  if (rpc < 1) rpc = 1;
  const int ray_ctas = (R + rpc - 1) / rpc;
  const int ray_grid = ray_ctas < sms ? ray_ctas : sms;
  const int head_grid = R < sms * 4 ? R : sms * 4;
  const int out_stride = ret_alpha ? 3 + S : 3;
  for (int i = 0; i < depth; ++i) {
    const float* lp = params + G_HEAD + (size_t)i * L_SIZE;
    if (precision != NFB_PREC_FP32) {
      float* ws = q + (size_t)N * D;
      float* attn = (ret_alpha && i == depth - 1) ? out + 3 : nullptr;
      rc = precision == NFB_PREC_BF16
               ? gnt_layer_tc<1>(i, R, S, V, ray_diff, mask, pts, ray_d, lp, F, q, ws, attn, out_stride, rpc, ray_grid, ray_block, sms, st)
               : gnt_layer_tc<3>(i, R, S, V, ray_diff, mask, pts, ray_d, lp, F, q, ws, attn, out_stride, rpc, ray_grid, ray_block, sms, st);
      if (rc) return rc;
      continue;
    }
    k_gnt_view_attn<<<grid_for(3), 128, sm_view, st>>>(N, V, F, ray_diff, mask, lp, q, q);
    NFB_CHECK_LAUNCH("k_gnt_view_attn");
    k_gnt_ffn<<<ffn_grid, 256, sm_ffn, st>>>(N, lp + L_V_LN2_W, q, q);
    NFB_CHECK_LAUNCH("k_gnt_ffn<view>");
    if ((i & 1) == 0) {
      k_gnt_qfc<<<grid_for(3), 128, sm_qfc, st>>>(N, S, pts, ray_d, lp, q, q);
      NFB_CHECK_LAUNCH("k_gnt_qfc");
    }
    const bool last = ret_alpha && i == depth - 1;
    k_gnt_ray_attn<<<ray_grid, ray_block * rpc, (size_t)(RS_W_TOTAL + (size_t)rpc * (2 * S * D + RS_PER_RAY)) * sizeof(float), st>>>(
        R, S, rpc, lp, q, q, last ? out + 3 : nullptr, out_stride);
    NFB_CHECK_LAUNCH("k_gnt_ray_attn");
    k_gnt_ffn<<<ffn_grid, 256, sm_ffn, st>>>(N, lp + L_R_LN2_W, q, q);
    NFB_CHECK_LAUNCH("k_gnt_ffn<ray>");
  }
  k_gnt_head<<<head_grid, ray_block < 64 ? 64 : ray_block, sm_head, st>>>(R, S, params + G_HEAD + (size_t)depth * L_SIZE, q, out, out_stride);
  NFB_CHECK_LAUNCH("k_gnt_head");
  return NFB_OK;
}


// fp32 forward with checkpoints (the first half of nfb_gnt_bwd, nfb_gnt_bwd.cu)
int nfbgnt::gnt_forward_checkpoints(int R, int S, int V, int depth, const float* rgb_feat, const float* ray_diff, const float* mask,
                                    const float* pts, const float* ray_d, const float* params, float* F, float* CK, float* VPA,
                                    float* out, int ret_alpha, float* scratch, cudaStream_t st) {
  const int N = R * S;
  const size_t rows = (size_t)N * V, NB = (size_t)N * D;
  auto ck = [&](int i, int j) { return CK + NB * (size_t)(5 * i + j); };
  const int sms = nfb_num_sms();
  int rc;
  const size_t sm_embed = (size_t)(35 * D + D + D * D + D) * sizeof(float), sm_view = (size_t)VS_TOTAL * sizeof(float),
               sm_ffn = (size_t)FS_TOTAL * sizeof(float), sm_qfc = (size_t)QS_TOTAL * sizeof(float);
  const int ray_block = ((S + 31) / 32) * 32;
  int rpc = 256 / ray_block;
  const int rpc_smem = (int)((200 * 1024 - (size_t)RS_W_TOTAL * sizeof(float)) / ((size_t)(2 * S * D + RS_PER_RAY) * sizeof(float)));
  if (rpc > rpc_smem) rpc = rpc_smem;
  if (rpc < 1) rpc = 1;
  const size_t sm_ray = (size_t)(RS_W_TOTAL + (size_t)rpc * (2 * S * D + RS_PER_RAY)) * sizeof(float);
  if ((rc = set_smem(k_gnt_embed, sm_embed, "k_gnt_embed"))) return rc;
  if ((rc = set_smem(k_gnt_view_attn, sm_view, "k_gnt_view_attn"))) return rc;
  if ((rc = set_smem(k_gnt_ffn, sm_ffn, "k_gnt_ffn"))) return rc;
  if ((rc = set_smem(k_gnt_qfc, sm_qfc, "k_gnt_qfc"))) return rc;
  if ((rc = set_smem(k_gnt_ray_attn, sm_ray, "k_gnt_ray_attn"))) return rc;
  const size_t sm_pre = (size_t)(2 * D + D * D) * sizeof(float);
  if ((rc = set_smem(k_gnt_view_row_fwd, sm_view, "k_gnt_view_row_fwd"))) return rc;
  auto grid_n = [&](size_t n, int block, int per_sm) {
    size_t g = (n + block - 1) / block;
    if (g > (size_t)sms * per_sm) g = (size_t)sms * per_sm;
    return (int)(g < 1 ? 1 : g);
  };
  const int ffn_grid = grid_n(N, 256, 1);
  const int ray_ctas = (R + rpc - 1) / rpc, ray_grid = ray_ctas < sms ? ray_ctas : sms;
  k_gnt_embed<<<grid_n(rows, 128, 8), 128, sm_embed, st>>>(rows, rgb_feat, params, F);
  NFB_CHECK_LAUNCH("k_gnt_embed");
  k_gnt_qinit<<<grid_n(NB, 256, 16), 256, 0, st>>>(NB, V, F, ck(0, 0));
  NFB_CHECK_LAUNCH("k_gnt_qinit");
  for (int i = 0; i < depth; ++i) {
    const float* lp = params + G_HEAD + (size_t)i * L_SIZE;
    float* vp_i = VPA + (size_t)i * rows * (D + 8);           // layer i: VP [rows][64] then A8 [rows][8]
    if (scratch) {
      float* qq = scratch;                                    // [N][64]
      float* xa = scratch + NB;                               // [N][64]
      k_gnt_pre<<<grid_n(N, 128, 6), 128, sm_pre, st>>>(N, lp, ck(i, 0), qq);
      NFB_CHECK_LAUNCH("k_gnt_pre");
      k_gnt_view_row_fwd<<<grid_n(rows, 128, 4), 128, sm_view, st>>>(rows, V, F, qq, ray_diff, lp, vp_i, vp_i + rows * D);
      NFB_CHECK_LAUNCH("k_gnt_view_row_fwd");
      int g = (N + 31) / 32;                                  // 8 threads per sample, 256 per CTA
      if (g > sms * 8) g = sms * 8;
      k_gnt_view_core<<<g, 256, 0, st>>>(N, V, vp_i + rows * D, vp_i, mask, lp, xa);
      NFB_CHECK_LAUNCH("k_gnt_view_core");
      k_gnt_outfc<<<grid_n(N, 128, 6), 128, sm_pre, st>>>(N, lp + L_V_O_W, lp + L_V_O_B, xa, ck(i, 0), ck(i, 1));
      NFB_CHECK_LAUNCH("k_gnt_outfc");
    } else {
      k_gnt_view_attn<<<grid_n(N, 128, 3), 128, sm_view, st>>>(N, V, F, ray_diff, mask, lp, ck(i, 0), ck(i, 1), vp_i, vp_i + rows * D);
      NFB_CHECK_LAUNCH("k_gnt_view_attn");
    }
    k_gnt_ffn<<<ffn_grid, 256, sm_ffn, st>>>(N, lp + L_V_LN2_W, ck(i, 1), ck(i, 2));
    NFB_CHECK_LAUNCH("k_gnt_ffn<view>");
    const float* qd = ck(i, 2);
    if ((i & 1) == 0) {
      k_gnt_qfc<<<grid_n(N, 128, 3), 128, sm_qfc, st>>>(N, S, pts, ray_d, lp, ck(i, 2), ck(i, 3));
      NFB_CHECK_LAUNCH("k_gnt_qfc");
      qd = ck(i, 3);
    }
    const int out_stride = ret_alpha ? 3 + S : 3;
    float* attn = (out && ret_alpha && i == depth - 1) ? out + 3 : nullptr;      // query 0's attention row of the last layer (:200)
    k_gnt_ray_attn<<<ray_grid, ray_block * rpc, sm_ray, st>>>(R, S, rpc, lp, qd, ck(i, 4), attn, out_stride);
    NFB_CHECK_LAUNCH("k_gnt_ray_attn");
    k_gnt_ffn<<<ffn_grid, 256, sm_ffn, st>>>(N, lp + L_R_LN2_W, ck(i, 4), ck(i, 5));
    NFB_CHECK_LAUNCH("k_gnt_ffn<ray>");
  }
  if (out) {
    const size_t sm_head = (size_t)(S * 65 + D) * sizeof(float);
    if ((rc = set_smem(k_gnt_head, sm_head, "k_gnt_head"))) return rc;
    const int head_grid = R < sms * 4 ? R : sms * 4;
    k_gnt_head<<<head_grid, ray_block < 64 ? 64 : ray_block, sm_head, st>>>(R, S, params + G_HEAD + (size_t)depth * L_SIZE, ck(depth, 0), out,
                                                                         ret_alpha ? 3 + S : 3);
    NFB_CHECK_LAUNCH("k_gnt_head");
  }
  return NFB_OK;
}
