// Shared definitions of the GNT kernels (nfb_gnt.cu: forward, nfb_gnt_bwd.cu: data gradient): parameter-blob layout,
// shared-memory layouts of the fp32 kernels, small row helpers.
#pragma once
#include "nfb_dense.cuh"

namespace nfbgnt {

constexpr int D = 64;          // netwidth (eval/gnt/config.py:111)
constexpr int DH = 256;        // feed-forward hidden width (4 x netwidth)
constexpr int PE = 63;         // 3 + 3 * 2 * 10 positional-encoding width (transformer_network.py:253-268)
constexpr int QIN = D + 2 * PE;  // 190

// ---- parameter blob layout (floats; every tensor in torch's [out][in] layout) ------------------------------
enum : int {
  G_RF0_W = 0,                       // rgbfeat_fc.0.weight [64][35]
  G_RF0_B = G_RF0_W + D * 35,
  G_RF2_W = G_RF0_B + D,             // rgbfeat_fc.2.weight [64][64]
  G_RF2_B = G_RF2_W + D * D,
  G_HEAD = G_RF2_B + D               // end of the header block
};
// per-layer block
enum : int {
  L_V_LN1_W = 0,                     // view_crosstrans.i.attn_norm
  L_V_LN1_B = L_V_LN1_W + D,
  L_V_Q = L_V_LN1_B + D,             // attn.q_fc / k_fc / v_fc .weight [64][64]
  L_V_K = L_V_Q + D * D,
  L_V_V = L_V_K + D * D,
  L_V_POS0_W = L_V_V + D * D,        // attn.pos_fc.0 [8][4]
  L_V_POS0_B = L_V_POS0_W + 32,
  L_V_POS2_W = L_V_POS0_B + 8,       // attn.pos_fc.2 [64][8]
  L_V_POS2_B = L_V_POS2_W + D * 8,
  L_V_AT0_W = L_V_POS2_B + D,        // attn.attn_fc.0 [8][64]
  L_V_AT0_B = L_V_AT0_W + 8 * D,
  L_V_AT2_W = L_V_AT0_B + 8,         // attn.attn_fc.2 [64][8]
  L_V_AT2_B = L_V_AT2_W + D * 8,
  L_V_O_W = L_V_AT2_B + D,           // attn.out_fc [64][64] + bias
  L_V_O_B = L_V_O_W + D * D,
  L_V_LN2_W = L_V_O_B + D,           // ff_norm
  L_V_LN2_B = L_V_LN2_W + D,
  L_V_FF1_W = L_V_LN2_B + D,         // ff.fc1 [256][64]
  L_V_FF1_B = L_V_FF1_W + DH * D,
  L_V_FF2_W = L_V_FF1_B + DH,        // ff.fc2 [64][256]
  L_V_FF2_B = L_V_FF2_W + D * DH,
  L_Q0_W = L_V_FF2_B + D,            // q_fcs.i.0 [64][190]   (even layers; the slot is unused on odd layers)
  L_Q0_B = L_Q0_W + D * QIN,
  L_Q2_W = L_Q0_B + D,               // q_fcs.i.2 [64][64]
  L_Q2_B = L_Q2_W + D * D,
  L_R_LN1_W = L_Q2_B + D,            // view_selftrans.i.attn_norm
  L_R_LN1_B = L_R_LN1_W + D,
  L_R_Q = L_R_LN1_B + D,             // attn.q_fc / k_fc / v_fc [64][64]
  L_R_K = L_R_Q + D * D,
  L_R_V = L_R_K + D * D,
  L_R_O_W = L_R_V + D * D,           // attn.out_fc + bias
  L_R_O_B = L_R_O_W + D * D,
  L_R_LN2_W = L_R_O_B + D,           // ff_norm
  L_R_LN2_B = L_R_LN2_W + D,
  L_R_FF1_W = L_R_LN2_B + D,
  L_R_FF1_B = L_R_FF1_W + DH * D,
  L_R_FF2_W = L_R_FF1_B + DH,
  L_R_FF2_B = L_R_FF2_W + D * DH,
  L_SIZE = L_R_FF2_B + D
};
// tail block (after depth layers): norm.weight, norm.bias, rgb_fc.weight [3][64], rgb_fc.bias [3]
enum : int { T_LN_W = 0, T_LN_B = D, T_RGB_W = 2 * D, T_RGB_B = 2 * D + 3 * D, T_SIZE = 2 * D + 3 * D + 3 };

constexpr float LN_EPS_T = 1e-6f;   // Transformer / Transformer2D norms (transformer_network.py:96-97,182-183)
constexpr float LN_EPS_HEAD = 1e-5f;  // GNT.norm = nn.LayerNorm default (:250)

__device__ __forceinline__ void layer_norm64(const float (&x)[D], const float* __restrict__ w, const float* __restrict__ b,
                                             float eps, float (&y)[D]) {
  float mu = 0.f;
#pragma unroll
  for (int c = 0; c < D; ++c) mu += x[c];
  mu *= (1.f / D);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < D; ++c) var = fmaf(x[c] - mu, x[c] - mu, var);
  var *= (1.f / D);
  const float rstd = 1.f / sqrtf(var + eps);
#pragma unroll
  for (int c = 0; c < D; ++c) y[c] = fmaf((x[c] - mu) * rstd, w[c], b[c]);
}

__device__ __forceinline__ void load_row64(const float* __restrict__ p, float (&x)[D]) {
#pragma unroll
  for (int c = 0; c < D; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + c);
    x[c] = t.x; x[c + 1] = t.y; x[c + 2] = t.z; x[c + 3] = t.w;
  }
}
__device__ __forceinline__ void store_row64(float* __restrict__ p, const float (&x)[D]) {
#pragma unroll
  for (int c = 0; c < D; c += 4) *reinterpret_cast<float4*>(p + c) = make_float4(x[c], x[c + 1], x[c + 2], x[c + 3]);
}

// shared-memory layouts of the fp32 forward kernels (the backward re-computes with the same layouts and code)
enum : int {
  VS_LN_W = 0, VS_LN_B = D, VS_Q = 2 * D, VS_K = VS_Q + D * D, VS_V = VS_K + D * D, VS_O = VS_V + D * D,
  VS_O_B = VS_O + D * D, VS_P0 = VS_O_B + D /*[4][8]*/, VS_P0_B = VS_P0 + 32, VS_P2 = VS_P0_B + 8 /*[8][64]*/,
  VS_P2_B = VS_P2 + 8 * D, VS_A0 = VS_P2_B + D /*[64][8]*/, VS_A0_B = VS_A0 + D * 8, VS_A2 = VS_A0_B + 8 /*[8][64]*/,
  VS_A2_B = VS_A2 + 8 * D, VS_TOTAL = VS_A2_B + D
};
enum : int { FS_LN_W = 0, FS_LN_B = D, FS_W1 = 2 * D /*[64][256]*/, FS_B1 = FS_W1 + D * DH, FS_W2 = FS_B1 + DH /*[256][64]*/,
             FS_B2 = FS_W2 + DH * D, FS_TOTAL = FS_B2 + D };
enum : int { QS_W0 = 0 /*[190][64]*/, QS_B0 = QIN * D, QS_W2 = QS_B0 + D, QS_B2 = QS_W2 + D * D, QS_TOTAL = QS_B2 + D };
enum : int { RS_LN_W = 0, RS_LN_B = D, RS_Q = 2 * D, RS_K = RS_Q + D * D, RS_V = RS_K + D * D, RS_O = RS_V + D * D,
             RS_O_B = RS_O + D * D, RS_W_TOTAL = RS_O_B + D,
             // per ray of the CTA: query 0 (scaled) [64], its softmax statistics m[4], 1/l[4]; then K [S][64], V [S][64]
             RS_Q0 = 0, RS_ST = D, RS_PER_RAY = D + 8 };

__device__ __forceinline__ void posenc_axpy(float (&h)[D], const float (&x3)[3], const float* __restrict__ w /*[63][64]*/) {
#pragma unroll
  for (int i = 0; i < 3; ++i) axpy_row<D>(h, x3[i], w + i * D);
#pragma unroll 1
  for (int f = 0; f < 10; ++f) {
    const float fr = (float)(1 << f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float ang = __fmul_rn(x3[i], fr);
      axpy_row<D>(h, sinf(ang), w + (3 + 6 * f + i) * D);
      axpy_row<D>(h, cosf(ang), w + (3 + 6 * f + 3 + i) * D);
    }
  }
}

// fp32 forward on the CUDA cores writing the running query after every block into ck[5 * i + j] (nfb_gnt.cu); F [rows][64],
// the 5 * depth + 1 checkpoints [N][64] and, per layer, the view attention's per-row v + pos [rows][64] | ReLU(attn_fc.0) [rows][8]
// (VPA: depth x rows x 72 floats) are the caller's; out != NULL: also the network output [R][3 (+ S)] (head + attention row of query 0)
int gnt_forward_checkpoints(int R, int S, int V, int depth, const float* rgb_feat, const float* ray_diff, const float* mask,
                            const float* pts, const float* ray_d, const float* params, float* F, float* CK, float* VPA,
                            float* out, int ret_alpha, float* scratch /* 2 x [N][64] or NULL: fused view attention */, cudaStream_t st);

}  // namespace nfbgnt
