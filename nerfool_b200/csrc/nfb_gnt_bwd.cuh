// GNT data gradient: d loss / d rgb_feat (and d ray_diff) of GNT.forward
// (gnt/transformer_network.py:270-309), what eval/gnt/eval_adv.py:282-545 back-propagates to the source-image
// perturbation through Projector.compute.  Included by nfb_gnt.cu inside its anonymous namespace (it uses the parameter
// layout and the fp32 forward kernels defined there).
//
// Shape of the computation (nfb_gnt_bwd):
//   1. checkpointing forward in fp32 on the CUDA cores: the forward kernels of nfb_gnt.cu write the running query after
//      every block into its own [N][64] buffer (5 per layer).  fp32, not the tensor-core forward: the network is full of
//      ReLUs, and a gradient is only comparable with the reference's if the ReLU masks are the reference's -- a
//      pre-activation that differs by 1e-5 relative flips ~1e-5 of the masks, fp32 arithmetic ~1e-7.
//   2. reverse sweep, every block re-computing its internals from its checkpoint with the SAME code as the forward
//      (bit-identical masks; the view attention's per-row products are saved by the forward instead, 288 B per row and layer):  head -> { FFN, ray attention, [q_fc], FFN, view attention } x depth -> max over views ->
//      rgbfeat_fc.  One thread per sample (or per (sample, view) row), weights in shared memory, like the forward.
//   The projected view features F are shared by all layers, so d F accumulates over the sweep (each row is owned by one
//   thread: plain read-modify-write, no atomics).
#pragma once

// LayerNorm statistics in the operation order of layer_norm64
__device__ __forceinline__ void ln_stats64(const float (&x)[D], float eps, float& mu, float& rstd) {
  mu = 0.f;
#pragma unroll
  for (int c = 0; c < D; ++c) mu += x[c];
  mu *= (1.f / D);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < D; ++c) var = fmaf(x[c] - mu, x[c] - mu, var);
  var *= (1.f / D);
  rstd = 1.f / sqrtf(var + eps);
}
// in place: d <- d LN(x) / d x applied to the cotangent d:  rstd (g - mean(g) - xhat mean(g xhat)),  g = d * w
__device__ __forceinline__ void layer_norm64_bwd(const float (&x)[D], const float* __restrict__ w, float eps, float (&d)[D]) {
  float mu, rstd;
  ln_stats64(x, eps, mu, rstd);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    d[c] *= w[c];
    s1 += d[c];
    s2 = fmaf(d[c], (x[c] - mu) * rstd, s2);
  }
  s1 *= (1.f / D);
  s2 *= (1.f / D);
#pragma unroll
  for (int c = 0; c < D; ++c) d[c] = rstd * (d[c] - s1 - (x[c] - mu) * rstd * s2);
}

__device__ __forceinline__ void load_nat(float* __restrict__ dst, const float* __restrict__ src, int n, int tid, int nt) {
  for (int i = tid; i < n; i += nt) dst[i] = __ldg(src + i);
}

// Warp-coalesced row I/O for the thread-per-row kernels: a thread's own 256-byte row is 16 requests that each touch 32 different
// lines across the warp; here the warp moves its 32 contiguous rows as 16 whole-line requests and the rows change hands through a
// per-warp staging block ws[32][CS] (CS = 68 floats: 272-byte stride, conflict-free 16-byte reads).  Every lane of the warp must call.
constexpr int CS = 68;
__device__ __forceinline__ void coop_load64(const float* __restrict__ base, size_t row0, size_t nrows, float* __restrict__ ws, int lane,
                                            bool act, float (&x)[D]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int piece = 32 * i + lane, r = piece >> 4, c = piece & 15;
    if (row0 + r < nrows) *reinterpret_cast<float4*>(ws + r * CS + 4 * c) = *(reinterpret_cast<const float4*>(base + (row0 + r) * D) + c);
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < D; c += 4) {
    const float4 t4 = act ? *reinterpret_cast<const float4*>(ws + lane * CS + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    x[c] = t4.x; x[c + 1] = t4.y; x[c + 2] = t4.z; x[c + 3] = t4.w;
  }
  __syncwarp();
}
__device__ __forceinline__ void coop_store64(float* __restrict__ base, size_t row0, size_t nrows, float* __restrict__ ws, int lane,
                                             const float (&y)[D]) {
#pragma unroll
  for (int c = 0; c < D; c += 4) *reinterpret_cast<float4*>(ws + lane * CS + c) = make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int piece = 32 * i + lane, r = piece >> 4, c = piece & 15;
    if (row0 + r < nrows) *(reinterpret_cast<float4*>(base + (row0 + r) * D) + c) = *reinterpret_cast<const float4*>(ws + r * CS + 4 * c);
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
// head (:303-305) backward: rgb = rgb_fc(mean_s LN(q)) -> dq[n] = LN'(q[n]) (rgb_fc^T d_rgb / S)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_gnt_head_bwd(int N, int S, const float* __restrict__ tp, const float* __restrict__ q,
                                                       const float* __restrict__ d_out, int out_stride, float* __restrict__ dq) {
  __shared__ __align__(16) float sw[D + 3 * D];
  load_nat(sw, tp + T_LN_W, D, threadIdx.x, blockDim.x);
  load_nat(sw + D, tp + T_RGB_W, 3 * D, threadIdx.x, blockDim.x);
  __syncthreads();
  const float inv_s = 1.f / (float)S;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    const int r = n / S;
    const float g0 = __ldg(d_out + (size_t)r * out_stride) * inv_s, g1 = __ldg(d_out + (size_t)r * out_stride + 1) * inv_s,
                g2 = __ldg(d_out + (size_t)r * out_stride + 2) * inv_s;
    float x[D], d[D];
    load_row64(q + (size_t)n * D, x);
#pragma unroll
    for (int c = 0; c < D; ++c) d[c] = fmaf(g0, sw[D + c], fmaf(g1, sw[2 * D + c], g2 * sw[3 * D + c]));
    layer_norm64_bwd(x, sw, LN_EPS_HEAD, d);
    store_row64(dq + (size_t)n * D, d);
  }
}

// ---------------------------------------------------------------------------------------------------
// feed-forward block backward: y = fc2(ReLU(fc1(LN(x)))) + x  ->  dq <- dq + LN'(fc1^T ((fc2^T dq) . [h > 0]))
// same shared-memory layout and the same hidden-unit arithmetic as k_gnt_ffn
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) k_gnt_ffn_bwd(int N, const float* __restrict__ lp, const float* __restrict__ q_in,
                                                         float* __restrict__ dq) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, nt = blockDim.x;          // nt == 256 (the dx columns below are sized for it)
  float* sdx = sm + FS_B2;                              // [64][256]: column t = this thread's fc1^T dh accumulator
  load_vec_padded(sm + FS_LN_W, lp, D, D, t, nt);
  load_vec_padded(sm + FS_LN_B, lp + D, D, D, t, nt);
  load_wt_transposed(sm + FS_W1, lp + 2 * D, DH, D, DH, t, nt);
  load_vec_padded(sm + FS_B1, lp + 2 * D + DH * D, DH, DH, t, nt);
  load_wt_transposed(sm + FS_W2, lp + 2 * D + DH * D + DH, D, DH, D, t, nt);
  __syncthreads();
  for (int n = blockIdx.x * blockDim.x + t; n < N; n += gridDim.x * blockDim.x) {
    float x[D], dy[D];
    {
      float q0[D];
      load_row64(q_in + (size_t)n * D, q0);
      layer_norm64(q0, sm + FS_LN_W, sm + FS_LN_B, LN_EPS_T, x);
    }
    load_row64(dq + (size_t)n * D, dy);
#pragma unroll 1
    for (int j0 = 0; j0 < DH; j0 += 32) {
      float h[32];
      load_bias<32>(h, sm + FS_B1 + j0);
#pragma unroll
      for (int k = 0; k < D; ++k) axpy_row<32>(h, x[k], sm + FS_W1 + k * DH + j0);
#pragma unroll 4
      for (int j = 0; j < 32; ++j) {
        const float g = dot_row<D>(dy, sm + FS_W2 + (j0 + j) * D);
        h[j] = h[j] > 0.f ? g : 0.f;
      }
#pragma unroll 4
      for (int k = 0; k < D; ++k) {
        const float s = dot_row<32>(h, sm + FS_W1 + k * DH + j0);
        sdx[k * 256 + t] = j0 == 0 ? s : sdx[k * 256 + t] + s;
      }
    }
    float dx[D];
#pragma unroll
    for (int k = 0; k < D; ++k) dx[k] = sdx[k * 256 + t];
    {
      float q0[D];
      load_row64(q_in + (size_t)n * D, q0);
      layer_norm64_bwd(q0, sm + FS_LN_W, LN_EPS_T, dx);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) dx[c] += dy[c];
    store_row64(dq + (size_t)n * D, dx);
  }
}

// ---------------------------------------------------------------------------------------------------
// projections in front of an attention core, per sample:  x = LN(q_in);  y_i = s_i W_i x  (i < nw, the forward's
// operation order);  g = Wo^T dy  (out_fc transposed applied to the block's cotangent)
// ---------------------------------------------------------------------------------------------------
struct ProjArgs {
  int N, nw;
  const float* q_in; const float* dy;
  const float* ln_w; const float* ln_b;
  const float* w[3]; float s0;
  const float* wo;
  float* y[3]; float* g;
};
__global__ void __launch_bounds__(128) k_gnt_proj(ProjArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* s_ln = sm;                  // 128
  float* s_w = sm + 2 * D;           // nw x [64][64] transposed
  float* s_o = s_w + 3 * D * D;      // [64][64] natural
  const int t = threadIdx.x, nt = blockDim.x;
  load_nat(s_ln, a.ln_w, D, t, nt);
  load_nat(s_ln + D, a.ln_b, D, t, nt);
  for (int i = 0; i < a.nw; ++i) load_wt_transposed(s_w + i * D * D, a.w[i], D, D, D, t, nt);
  load_nat(s_o, a.wo, D * D, t, nt);
  __syncthreads();
  for (int n = blockIdx.x * blockDim.x + t; n < a.N; n += gridDim.x * blockDim.x) {
    {
      float x[D];
      {
        float q0[D];
        load_row64(a.q_in + (size_t)n * D, q0);
        layer_norm64(q0, s_ln, s_ln + D, LN_EPS_T, x);
      }
      for (int i = 0; i < a.nw; ++i) {
        float y[D];
#pragma unroll
        for (int c = 0; c < D; ++c) y[c] = 0.f;
        dense_acc<D, D>(s_w + i * D * D, x, y);
        if (i == 0) {
#pragma unroll
          for (int c = 0; c < D; ++c) y[c] *= a.s0;
        }
        store_row64(a.y[i] + (size_t)n * D, y);
      }
    }
    float dy[D], g[D];
    load_row64(a.dy + (size_t)n * D, dy);
#pragma unroll
    for (int c = 0; c < D; ++c) g[c] = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) axpy_row<D>(g, dy[k], s_o + k * D);
    store_row64(a.g + (size_t)n * D, g);
  }
}

// ---------------------------------------------------------------------------------------------------
// after an attention core, per sample:  dx = sum_i s_i W_i^T g_i  (i < nw)  ->  dq <- dq + LN'(dx)
// view attention (nv > 0): the single cotangent is g_0 = - sum_v DT[n][v]  (attn = k - q + pos, :80)
// ---------------------------------------------------------------------------------------------------
struct PostArgs {
  int N, nw, nv;
  const float* q_in; float* dq;
  const float* ln_w;
  const float* w[3]; float s0;
  const float* g[3];
};
__global__ void __launch_bounds__(128) k_gnt_post(PostArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* s_ln = sm;                  // 64
  float* s_w = sm + D;               // nw x [64][64] natural
  const int t = threadIdx.x, nt = blockDim.x;
  load_nat(s_ln, a.ln_w, D, t, nt);
  for (int i = 0; i < a.nw; ++i) load_nat(s_w + i * D * D, a.w[i], D * D, t, nt);
  __syncthreads();
  for (int n = blockIdx.x * blockDim.x + t; n < a.N; n += gridDim.x * blockDim.x) {
    float dx[D];
#pragma unroll
    for (int c = 0; c < D; ++c) dx[c] = 0.f;
    for (int i = 0; i < a.nw; ++i) {
      float g[D];
      if (a.nv > 0) {
#pragma unroll
        for (int c = 0; c < D; ++c) g[c] = 0.f;
        for (int v = 0; v < a.nv; ++v) {
          float tv[D];
          load_row64(a.g[0] + ((size_t)n * a.nv + v) * D, tv);
#pragma unroll
          for (int c = 0; c < D; ++c) g[c] -= tv[c];
        }
      } else {
        load_row64(a.g[i] + (size_t)n * D, g);
      }
      const float sc = i == 0 ? a.s0 : 1.f;
#pragma unroll
      for (int k = 0; k < D; ++k) axpy_row<D>(dx, g[k] * sc, s_w + i * D * D + k * D);
    }
    {
      float q0[D];
      load_row64(a.q_in + (size_t)n * D, q0);
      layer_norm64_bwd(q0, s_ln, LN_EPS_T, dx);
    }
    float dy[D];
    load_row64(a.dq + (size_t)n * D, dy);
#pragma unroll
    for (int c = 0; c < D; ++c) dx[c] += dy[c];
    store_row64(a.dq + (size_t)n * D, dx);
  }
}

// ---------------------------------------------------------------------------------------------------
// view attention core backward.  Two adjacent lanes share a sample, each owns 32 of the 64 channels (as k_gnt_view_core).
//   forward:  a_v = attn_fc.2(A8_v) (masked_fill -1e9), w_v = softmax_v(a_v) per channel, out = sum_v VP_v w_v
//   backward: dVP_v = g w_v ;  d a_v = w_v g (VP_v - out)  (0 for a masked row: masked_fill blocks it) ;
//             dA8_v = (attn_fc.2^T d a_v) . [A8_v > 0]          g = out_fc^T dy (from k_gnt_proj)
// VP / A8 are the rows the checkpointing forward (k_gnt_view_attn) saved for this layer; dVP / dA8 overwrite them in place.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_gnt_view_core_bwd(int N, int V, float* __restrict__ A8, float* __restrict__ VP,
                                                            const float* __restrict__ mask, const float* __restrict__ G,
                                                            const float* __restrict__ lp) {
  constexpr int HC = D / 2;
  __shared__ __align__(16) float sm[8 * D + D + D * 8];     // attn_fc.2 transposed [8][64], bias [64], natural [64][8]
  const int t = threadIdx.x, nt = blockDim.x;
  load_wt_transposed(sm, lp + L_V_AT2_W, D, 8, D, t, nt);
  load_vec_padded(sm + 8 * D, lp + L_V_AT2_B, D, D, t, nt);
  load_nat(sm + 9 * D, lp + L_V_AT2_W, D * 8, t, nt);
  __syncthreads();
  const int c0 = (t & 1) * HC;
  const int npairs = blockDim.x / 2;
  const int iters = (N + gridDim.x * npairs - 1) / (gridDim.x * npairs);
  for (int it = 0; it < iters; ++it) {
    const int n = (it * gridDim.x + blockIdx.x) * npairs + (t >> 1);
    const bool act = n < N;                                  // both lanes of a pair agree; idle pairs still run the shuffles
    const int ns = act ? n : 0;
    float m[HC], l[HC], acc[HC];
#pragma unroll
    for (int c = 0; c < HC; ++c) { m[c] = -3.4e38f; l[c] = 0.f; acc[c] = 0.f; }
    for (int v = 0; v < V; ++v) {
      const size_t row = (size_t)ns * V + v;
      const float4 h0 = *reinterpret_cast<const float4*>(A8 + row * 8), h1 = *(reinterpret_cast<const float4*>(A8 + row * 8) + 1);
      const float a8[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
      const bool valid = __ldg(mask + row) != 0.f;
      const float4* vr = reinterpret_cast<const float4*>(VP + row * D + c0);
#pragma unroll
      for (int cc = 0; cc < HC; cc += 16) {
        float a[16];
        load_bias<16>(a, sm + 8 * D + c0 + cc);
#pragma unroll
        for (int j = 0; j < 8; ++j) axpy_row<16>(a, a8[j], sm + j * D + c0 + cc);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v4 = vr[cc / 4 + j];
          const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = cc + 4 * j + i;
            const float s = valid ? a[4 * j + i] : -1e9f;
            const float mn = fmaxf(m[c], s);
            const float sc = __expf(m[c] - mn), e = __expf(s - mn);
            l[c] = fmaf(l[c], sc, e);
            acc[c] = fmaf(acc[c], sc, vv[i] * e);
            m[c] = mn;
          }
        }
      }
    }
    float g[HC];
#pragma unroll
    for (int c = 0; c < HC; c += 4) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(G + (size_t)ns * D + c0 + c));
      g[c] = g4.x; g[c + 1] = g4.y; g[c + 2] = g4.z; g[c + 3] = g4.w;
      l[c] = 1.f / l[c]; l[c + 1] = 1.f / l[c + 1]; l[c + 2] = 1.f / l[c + 2]; l[c + 3] = 1.f / l[c + 3];
      acc[c] *= l[c]; acc[c + 1] *= l[c + 1]; acc[c + 2] *= l[c + 2]; acc[c + 3] *= l[c + 3];       // out
    }
    for (int v = 0; v < V; ++v) {
      const size_t row = (size_t)ns * V + v;
      const float4 h0 = *reinterpret_cast<const float4*>(A8 + row * 8), h1 = *(reinterpret_cast<const float4*>(A8 + row * 8) + 1);
      const float a8[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
      const bool valid = __ldg(mask + row) != 0.f;
      float4* vr = reinterpret_cast<float4*>(VP + row * D + c0);
      float d8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) d8[j] = 0.f;
#pragma unroll
      for (int cc = 0; cc < HC; cc += 16) {
        float a[16];
        load_bias<16>(a, sm + 8 * D + c0 + cc);
#pragma unroll
        for (int j = 0; j < 8; ++j) axpy_row<16>(a, a8[j], sm + j * D + c0 + cc);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v4 = vr[cc / 4 + j];
          const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
          float dv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = cc + 4 * j + i;
            const float s = valid ? a[4 * j + i] : -1e9f;
            const float w = __expf(s - m[c]) * l[c];
            dv[i] = g[c] * w;
            const float da = valid ? dv[i] * (vv[i] - acc[c]) : 0.f;
            const float* w2 = sm + 9 * D + (c0 + c) * 8;             // attn_fc.2.weight[c][0..8)
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) d8[jj] = fmaf(da, w2[jj], d8[jj]);
          }
          if (act) vr[cc / 4 + j] = make_float4(dv[0], dv[1], dv[2], dv[3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        d8[j] += __shfl_xor_sync(0xffffffffu, d8[j], 1);
        d8[j] = a8[j] > 0.f ? d8[j] : 0.f;
      }
      if (act && (t & 1) == 0) {
        float4* o8 = reinterpret_cast<float4*>(A8 + row * 8);
        o8[0] = make_float4(d8[0], d8[1], d8[2], d8[3]);
        o8[1] = make_float4(d8[4], d8[5], d8[6], d8[7]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// view attention, per (sample, view) row, backward of the row-wise part:
//   dt = attn_fc.0^T dA8 ;  dk = dt + v_fc^T dVP ;  dF += k_fc^T dk ;  DT = dt (summed over the views by k_gnt_post: d qq = -sum)
//   optional:  d ray_diff += pos_fc.0^T ((pos_fc.2^T (dt + dVP)) . [p8 > 0])
// DT overwrites the dVP row.
// ---------------------------------------------------------------------------------------------------
enum : int { VB_A0 = 0 /*[8][64] natural*/, VB_V = VB_A0 + 8 * D /*[64][64] natural*/, VB_K = VB_V + D * D,
             VB_P2 = VB_K + D * D /*[64][8] natural*/, VB_P0 = VB_P2 + D * 8 /*[8][4] natural*/, VB_P0T = VB_P0 + 32 /*[4][8]*/,
             VB_P0_B = VB_P0T + 32, VB_TOTAL = VB_P0_B + 8 };
__global__ void __launch_bounds__(128) k_gnt_view_row_bwd(size_t rows, const float* __restrict__ DA8, float* __restrict__ DVP,
                                                           const float* __restrict__ ray_diff, const float* __restrict__ lp,
                                                           float* __restrict__ dF, float* __restrict__ d_ray_diff) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, nt = blockDim.x;
  load_nat(sm + VB_A0, lp + L_V_AT0_W, 8 * D, t, nt);
  load_nat(sm + VB_V, lp + L_V_V, D * D, t, nt);
  load_nat(sm + VB_K, lp + L_V_K, D * D, t, nt);
  load_nat(sm + VB_P2, lp + L_V_POS2_W, D * 8, t, nt);
  load_nat(sm + VB_P0, lp + L_V_POS0_W, 32, t, nt);
  load_wt_transposed(sm + VB_P0T, lp + L_V_POS0_W, 8, 4, 8, t, nt);
  load_vec_padded(sm + VB_P0_B, lp + L_V_POS0_B, 8, 8, t, nt);
  __syncthreads();
  const int lane = t & 31;
  float* ws = sm + VB_TOTAL + (t >> 5) * 32 * CS;            // this warp's staging block
  for (size_t rbase = (size_t)blockIdx.x * blockDim.x; rbase < rows; rbase += (size_t)gridDim.x * blockDim.x) {
    const size_t row = rbase + t, row0 = row - lane;
    const bool act = row < rows;
    const size_t rs = act ? row : 0;
    float dt[D];
    {
      const float4 h0 = __ldg(reinterpret_cast<const float4*>(DA8 + rs * 8)), h1 = __ldg(reinterpret_cast<const float4*>(DA8 + rs * 8) + 1);
      const float d8[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int c = 0; c < D; ++c) dt[c] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) axpy_row<D>(dt, d8[j], sm + VB_A0 + j * D);
    }
    float dk[D];
    {
      float dvp[D];
      coop_load64(DVP, row0, rows, ws, lane, act, dvp);
      coop_store64(DVP, row0, rows, ws, lane, dt);          // DT
      if (d_ray_diff && act) {
        float dp8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dp8[j] = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) axpy_row<8>(dp8, dt[c] + dvp[c], sm + VB_P2 + c * 8);
        const float4 rd4 = __ldg(reinterpret_cast<const float4*>(ray_diff) + row);
        const float rd[4] = {rd4.x, rd4.y, rd4.z, rd4.w};
        float p8[8];
        load_bias<8>(p8, sm + VB_P0_B);
        dense_acc<4, 8>(sm + VB_P0T, rd, p8);
        float dr[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (p8[j] > 0.f) axpy_row<4>(dr, dp8[j], sm + VB_P0 + j * 4);
        float4* o = reinterpret_cast<float4*>(d_ray_diff) + row;
        const float4 old = *o;
        *o = make_float4(old.x + dr[0], old.y + dr[1], old.z + dr[2], old.w + dr[3]);
      }
#pragma unroll
      for (int c = 0; c < D; ++c) dk[c] = dt[c];
#pragma unroll
      for (int k = 0; k < D; ++k) axpy_row<D>(dk, dvp[k], sm + VB_V + k * D);
    }
    float df[D];
    coop_load64(dF, row0, rows, ws, lane, act, df);
#pragma unroll
    for (int k = 0; k < D; ++k) axpy_row<D>(df, dk[k], sm + VB_K + k * D);
    coop_store64(dF, row0, rows, ws, lane, df);
  }
}

// ---------------------------------------------------------------------------------------------------
// q_fc backward (even layers, :295-297, no residual):  dq <- fc.0[:, :64]^T ((fc.2^T dq) . [h > 0]);  h re-computed with the
// code of k_gnt_qfc
// ---------------------------------------------------------------------------------------------------
enum : int { QB_W0 = 0 /*[190][64] transposed*/, QB_B0 = QIN * D, QB_W2N = QB_B0 + D /*fc.2 natural*/, QB_W0N = QB_W2N + D * D /*fc.0[:, :64] natural*/,
             QB_TOTAL = QB_W0N + D * D };
__global__ void __launch_bounds__(128) k_gnt_qfc_bwd(int N, int S, const float* __restrict__ pts, const float* __restrict__ ray_d,
                                                      const float* __restrict__ lp, const float* __restrict__ q_in, float* __restrict__ dq) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, nt = blockDim.x;
  load_wt_transposed(sm + QB_W0, lp + L_Q0_W, D, QIN, D, t, nt);
  load_vec_padded(sm + QB_B0, lp + L_Q0_B, D, D, t, nt);
  load_nat(sm + QB_W2N, lp + L_Q2_W, D * D, t, nt);
  for (int i = t; i < D * D; i += nt) sm[QB_W0N + i] = __ldg(lp + L_Q0_W + (i >> 6) * QIN + (i & 63));
  __syncthreads();
  for (int n = blockIdx.x * blockDim.x + t; n < N; n += gridDim.x * blockDim.x) {
    float h[D];
    load_bias<D>(h, sm + QB_B0);
    {
      float q0[D];
      load_row64(q_in + (size_t)n * D, q0);
      dense_acc<D, D>(sm + QB_W0, q0, h);
    }
    const float p3[3] = {__ldg(pts + (size_t)n * 3), __ldg(pts + (size_t)n * 3 + 1), __ldg(pts + (size_t)n * 3 + 2)};
    posenc_axpy(h, p3, sm + QB_W0 + D * D);
    const int r = n / S;
    const float dx = __ldg(ray_d + (size_t)r * 3), dy_ = __ldg(ray_d + (size_t)r * 3 + 1), dz = __ldg(ray_d + (size_t)r * 3 + 2);
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy_, dy_)), __fmul_rn(dz, dz)));
    const float d3[3] = {__fdiv_rn(dx, nrm), __fdiv_rn(dy_, nrm), __fdiv_rn(dz, nrm)};
    posenc_axpy(h, d3, sm + QB_W0 + (D + PE) * D);
    float g[D];
    {
      float dy[D];
      load_row64(dq + (size_t)n * D, dy);
#pragma unroll
      for (int c = 0; c < D; ++c) g[c] = 0.f;
#pragma unroll
      for (int k = 0; k < D; ++k) axpy_row<D>(g, dy[k], sm + QB_W2N + k * D);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) g[c] = h[c] > 0.f ? g[c] : 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c) h[c] = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) axpy_row<D>(h, g[k], sm + QB_W0N + k * D);
    store_row64(dq + (size_t)n * D, h);
  }
}

// ---------------------------------------------------------------------------------------------------
// ray attention core backward.  One CTA per `rpc` rays, one thread per sample, one head at a time; the head's 16-wide slices
// of Qs (= q / sqrt(16)), K, V and G (= out_fc^T dy) of the ray sit in shared memory.
//   P = softmax_j(Qs K^T),  O = P V ;   dV_j = sum_i P_ij G_i ;  dP_ij = G_i . V_j (+ d_alpha_j / 4 for query 0 of the last
//   layer: the network's second output is mean_h P_0j, :200) ;  dS_ij = P_ij (dP_ij - sum_j P_ij dP_ij) ;
//   dQs_i = sum_j dS_ij K_j ;  dK_j = sum_i dS_ij Qs_i.
// Query-side pass (thread = query i) then key-side pass (thread = key j) over the same shared slices: no atomics.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) k_gnt_ray_core_bwd(int R, int S, int rpc, const float* __restrict__ Qs, const float* __restrict__ K,
                                                              const float* __restrict__ Vp, const float* __restrict__ G,
                                                              const float* __restrict__ d_alpha, int alpha_stride,
                                                              float* __restrict__ dQ, float* __restrict__ dK, float* __restrict__ dV) {
  extern __shared__ __align__(16) float sm[];
  const int nt = blockDim.x;
  const int rb = nt / rpc;
  const int lr = threadIdx.x / rb;
  const int t = threadIdx.x - lr * rb;
  float* base = sm + (size_t)lr * (size_t)(4 * 16 + 4) * S;
  float* sq = base;                       // [S][16]
  float* sk = sq + (size_t)S * 16;
  float* sv = sk + (size_t)S * 16;
  float* sg = sv + (size_t)S * 16;
  float* s_m = sg + (size_t)S * 16;       // per query: max, 1 / sum, sum_j P dP
  float* s_il = s_m + S;
  float* s_dd = s_il + S;
  float* s_a = s_dd + S;                  // d_alpha row / 4 (zeros without it)
  for (int r0 = blockIdx.x * rpc; r0 < R; r0 += gridDim.x * rpc) {
    const int r = r0 + lr;
    const bool act = (t < S) && (r < R);
    const int Sr = (r < R) ? S : 0;
    const size_t n = (size_t)(r < R ? r : 0) * S + (t < S ? t : 0);
    if (t < S) s_a[t] = (act && d_alpha) ? 0.25f * __ldg(d_alpha + (size_t)r * alpha_stride + t) : 0.f;
    const float ga = (t == 0) ? 1.f : 0.f;
#pragma unroll 1
    for (int h = 0; h < 4; ++h) {
      float q[16], g[16];
      {
        float kk[16], vv[16];
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          const float4 a4 = __ldg(reinterpret_cast<const float4*>(Qs + n * D + 16 * h + c));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(K + n * D + 16 * h + c));
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(Vp + n * D + 16 * h + c));
          const float4 d4 = __ldg(reinterpret_cast<const float4*>(G + n * D + 16 * h + c));
          q[c] = a4.x; q[c + 1] = a4.y; q[c + 2] = a4.z; q[c + 3] = a4.w;
          kk[c] = b4.x; kk[c + 1] = b4.y; kk[c + 2] = b4.z; kk[c + 3] = b4.w;
          vv[c] = c4.x; vv[c + 1] = c4.y; vv[c + 2] = c4.z; vv[c + 3] = c4.w;
          g[c] = d4.x; g[c + 1] = d4.y; g[c + 2] = d4.z; g[c + 3] = d4.w;
        }
        if (act) {
#pragma unroll
          for (int c = 0; c < 16; c += 4) {
            *reinterpret_cast<float4*>(sq + t * 16 + c) = make_float4(q[c], q[c + 1], q[c + 2], q[c + 3]);
            *reinterpret_cast<float4*>(sk + t * 16 + c) = make_float4(kk[c], kk[c + 1], kk[c + 2], kk[c + 3]);
            *reinterpret_cast<float4*>(sv + t * 16 + c) = make_float4(vv[c], vv[c + 1], vv[c + 2], vv[c + 3]);
            *reinterpret_cast<float4*>(sg + t * 16 + c) = make_float4(g[c], g[c + 1], g[c + 2], g[c + 3]);
          }
        }
      }
      __syncthreads();
      // ---- query side: statistics (max, sum, sum_j P dP), then dQs
      float mx = -3.4e38f;
      for (int j = 0; j < Sr; ++j) {
        const float* kj = sk + j * 16;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) s = fmaf(q[c], kj[c], s);
        mx = fmaxf(mx, s);
      }
      float l = 0.f, pd = 0.f;
      for (int j = 0; j < Sr; ++j) {
        const float* kj = sk + j * 16;
        const float* vj = sv + j * 16;
        float s = 0.f, dp = ga * s_a[j];
#pragma unroll
        for (int c = 0; c < 16; ++c) { s = fmaf(q[c], kj[c], s); dp = fmaf(g[c], vj[c], dp); }
        const float p = __expf(s - mx);
        l += p;
        pd = fmaf(p, dp, pd);
      }
      const float il = Sr > 0 ? 1.f / l : 0.f;
      const float dd = pd * il;
      float dq16[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) dq16[c] = 0.f;
      for (int j = 0; j < Sr; ++j) {
        const float* kj = sk + j * 16;
        const float* vj = sv + j * 16;
        float s = 0.f, dp = ga * s_a[j];
#pragma unroll
        for (int c = 0; c < 16; ++c) { s = fmaf(q[c], kj[c], s); dp = fmaf(g[c], vj[c], dp); }
        const float ds = __expf(s - mx) * il * (dp - dd);
#pragma unroll
        for (int c = 0; c < 16; ++c) dq16[c] = fmaf(ds, kj[c], dq16[c]);
      }
      if (act) {
        s_m[t] = mx; s_il[t] = il; s_dd[t] = dd;
#pragma unroll
        for (int c = 0; c < 16; c += 4)
          *reinterpret_cast<float4*>(dQ + n * D + 16 * h + c) = make_float4(dq16[c], dq16[c + 1], dq16[c + 2], dq16[c + 3]);
      }
      __syncthreads();
      // ---- key side: dK_j, dV_j  (this thread's key = its own sample: k, v slices re-read from shared memory)
      {
        float kk[16], vv[16], dk[16], dv[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          kk[c] = sk[(t < S ? t : 0) * 16 + c];
          vv[c] = sv[(t < S ? t : 0) * 16 + c];
          dk[c] = 0.f; dv[c] = 0.f;
        }
        const float aj = s_a[t < S ? t : 0];
        for (int i = 0; i < Sr; ++i) {
          const float* qi = sq + i * 16;
          const float* gi = sg + i * 16;
          float s = 0.f, dp = (i == 0) ? aj : 0.f;
#pragma unroll
          for (int c = 0; c < 16; ++c) { s = fmaf(qi[c], kk[c], s); dp = fmaf(gi[c], vv[c], dp); }
          const float p = __expf(s - s_m[i]) * s_il[i];
          const float ds = p * (dp - s_dd[i]);
#pragma unroll
          for (int c = 0; c < 16; ++c) { dv[c] = fmaf(p, gi[c], dv[c]); dk[c] = fmaf(ds, qi[c], dk[c]); }
        }
        if (act) {
#pragma unroll
          for (int c = 0; c < 16; c += 4) {
            *reinterpret_cast<float4*>(dK + n * D + 16 * h + c) = make_float4(dk[c], dk[c + 1], dk[c + 2], dk[c + 3]);
            *reinterpret_cast<float4*>(dV + n * D + 16 * h + c) = make_float4(dv[c], dv[c + 1], dv[c + 2], dv[c + 3]);
          }
        }
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// q = max over the views of F (:286) backward: the cotangent goes to the FIRST view that attains the maximum (torch.max)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gnt_qinit_bwd(size_t n_elems, int V, const float* __restrict__ F, const float* __restrict__ dq,
                                                        float* __restrict__ dF) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / D;
    const int c = (int)(i - n * D);
    float m = -3.4e38f;
    int arg = 0;
    for (int v = 0; v < V; ++v) {
      const float f = F[(n * V + v) * D + c];
      if (f > m) { m = f; arg = v; }
    }
    dF[(n * V + arg) * D + c] += dq[i];
  }
}

// ---------------------------------------------------------------------------------------------------
// rgbfeat_fc backward, per (sample, view) row: d rgb_feat = fc.0^T ((fc.2^T dF) . [h > 0]),  h as in k_gnt_embed
// ---------------------------------------------------------------------------------------------------
enum : int { EB_W0 = 0 /*[35][64] transposed*/, EB_B0 = 35 * D, EB_W2N = EB_B0 + D /*fc.2 natural*/, EB_W0N = EB_W2N + D * D /*fc.0 natural [64][36]*/,
             EB_TOTAL = EB_W0N + D * 36 };
__global__ void __launch_bounds__(128) k_gnt_embed_bwd(size_t rows, const float* __restrict__ rgb_feat, const float* __restrict__ params,
                                                        const float* __restrict__ dF, float* __restrict__ d_rgb_feat) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, nt = blockDim.x;
  load_wt_transposed(sm + EB_W0, params + G_RF0_W, D, 35, D, t, nt);
  load_vec_padded(sm + EB_B0, params + G_RF0_B, D, D, t, nt);
  load_nat(sm + EB_W2N, params + G_RF2_W, D * D, t, nt);
  for (int i = t; i < D * 36; i += nt) sm[EB_W0N + i] = (i % 36) < 35 ? __ldg(params + G_RF0_W + (i / 36) * 35 + (i % 36)) : 0.f;
  __syncthreads();
  for (size_t r = (size_t)blockIdx.x * blockDim.x + t; r < rows; r += (size_t)gridDim.x * blockDim.x) {
    float h[D];
    load_bias<D>(h, sm + EB_B0);
    const float* x = rgb_feat + r * NFB_ROW_CH;
#pragma unroll
    for (int k = 0; k < NFB_ROW_CH; ++k) axpy_row<D>(h, __ldg(x + k), sm + EB_W0 + k * D);
    float g[D];
    {
      float df[D];
      load_row64(dF + r * D, df);
#pragma unroll
      for (int c = 0; c < D; ++c) g[c] = 0.f;
#pragma unroll
      for (int k = 0; k < D; ++k) axpy_row<D>(g, df[k], sm + EB_W2N + k * D);
    }
    float dx[36];
#pragma unroll
    for (int c = 0; c < 36; ++c) dx[c] = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) axpy_row<36>(dx, h[k] > 0.f ? g[k] : 0.f, sm + EB_W0N + k * 36);
    float* o = d_rgb_feat + r * NFB_ROW_CH;
#pragma unroll
    for (int c = 0; c < NFB_ROW_CH; ++c) o[c] = dx[c];
  }
}
