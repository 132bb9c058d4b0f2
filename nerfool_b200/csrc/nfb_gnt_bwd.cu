// GNT data gradient: nfb_gnt_bwd (kernels in nfb_gnt_bwd.cuh; the checkpointing fp32 forward lives in nfb_gnt.cu).
#include "nfb_gnt_common.cuh"

using namespace nfbgnt;

namespace {
#include "nfb_gnt_bwd.cuh"

template <typename K>
int set_smem(K kern, size_t bytes, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
  return NFB_OK;
}
}  // namespace

// ---------------------------------------------------------------------------------------------------
// data gradient (nfb_gnt_bwd.cuh)
// ---------------------------------------------------------------------------------------------------
extern "C" size_t nfb_gnt_bwd_workspace_bytes(int R, int S, int V, int depth) {
  // F, dF [rows][64], per layer VP [rows][64] + A8 [rows][8]  |  5 * depth + 1 checkpoints, dq and seven per-sample buffers [N][64]
  if (R <= 0 || S < 1 || V < 1 || depth < 1) return 0;
  const size_t N = (size_t)R * S, rows = N * V;
  return (rows * (2 * D + (size_t)depth * (D + 8)) + N * D * (size_t)(5 * depth + 1 + 1 + 7)) * sizeof(float);
}

// mode 0: checkpointing forward + reverse sweep (nfb_gnt_bwd) | 1: checkpointing forward only, writes `out` (nfb_gnt_fwd_save) |
// 2: reverse sweep on a workspace filled by mode 1 (nfb_gnt_bwd_saved)
static int gnt_grad_impl(int mode, int R, int S, int V, int depth, int ret_alpha, const float* rgb_feat, const float* ray_diff,
                         const float* mask, const float* pts, const float* ray_d, const float* params, float* out, const float* d_out,
                         float* d_rgb_feat, float* d_ray_diff, void* workspace, size_t workspace_bytes, void* stream);

extern "C" int nfb_gnt_bwd(int R, int S, int V, int depth, int ret_alpha, const float* rgb_feat, const float* ray_diff,
                           const float* mask, const float* pts, const float* ray_d, const float* params, const float* d_out,
                           float* d_rgb_feat, float* d_ray_diff, void* workspace, size_t workspace_bytes, void* stream) {
  return gnt_grad_impl(0, R, S, V, depth, ret_alpha, rgb_feat, ray_diff, mask, pts, ray_d, params, nullptr, d_out, d_rgb_feat, d_ray_diff,
                       workspace, workspace_bytes, stream);
}
extern "C" int nfb_gnt_fwd_save(int R, int S, int V, int depth, int ret_alpha, const float* rgb_feat, const float* ray_diff,
                                const float* mask, const float* pts, const float* ray_d, const float* params, float* out,
                                void* workspace, size_t workspace_bytes, void* stream) {
  return gnt_grad_impl(1, R, S, V, depth, ret_alpha, rgb_feat, ray_diff, mask, pts, ray_d, params, out, nullptr, nullptr, nullptr,
                       workspace, workspace_bytes, stream);
}
extern "C" int nfb_gnt_bwd_saved(int R, int S, int V, int depth, int ret_alpha, const float* rgb_feat, const float* ray_diff,
                                 const float* mask, const float* pts, const float* ray_d, const float* params, const float* d_out,
                                 float* d_rgb_feat, float* d_ray_diff, void* workspace, size_t workspace_bytes, void* stream) {
  return gnt_grad_impl(2, R, S, V, depth, ret_alpha, rgb_feat, ray_diff, mask, pts, ray_d, params, nullptr, d_out, d_rgb_feat, d_ray_diff,
                       workspace, workspace_bytes, stream);
}

static int gnt_grad_impl(int mode, int R, int S, int V, int depth, int ret_alpha, const float* rgb_feat, const float* ray_diff,
                         const float* mask, const float* pts, const float* ray_d, const float* params, float* out, const float* d_out,
                         float* d_rgb_feat, float* d_ray_diff, void* workspace, size_t workspace_bytes, void* stream) {
  NFB_REQUIRE(R >= 0 && S >= 1 && V >= 1 && depth >= 1, NFB_EINVAL, "nfb_gnt_bwd: bad arguments (R=%d S=%d V=%d depth=%d)", R, S, V, depth);
  NFB_REQUIRE(S <= NFB_MAX_SAMPLES, NFB_EUNSUPPORTED, "nfb_gnt_bwd: S=%d > %d samples per ray", S, NFB_MAX_SAMPLES);
  NFB_REQUIRE(V <= NFB_MAX_VIEWS, NFB_EUNSUPPORTED, "nfb_gnt_bwd: V=%d > %d views", V, NFB_MAX_VIEWS);
  if (R == 0) return NFB_OK;
  NFB_REQUIRE(rgb_feat && ray_diff && mask && pts && ray_d && params && workspace && (mode == 1 ? out != nullptr : (d_out && d_rgb_feat)),
              NFB_EINVAL, "nfb_gnt_bwd: NULL buffer");
  NFB_REQUIRE(workspace_bytes >= nfb_gnt_bwd_workspace_bytes(R, S, V, depth), NFB_EINVAL, "nfb_gnt_bwd: workspace too small (%zu < %zu bytes)",
              workspace_bytes, nfb_gnt_bwd_workspace_bytes(R, S, V, depth));
  NFB_REQUIRE(((uintptr_t)ray_diff % 16) == 0 && ((uintptr_t)workspace % 16) == 0 && ((uintptr_t)params % 16) == 0 &&
                  ((uintptr_t)d_ray_diff % 16) == 0, NFB_EINVAL, "nfb_gnt_bwd: ray_diff / d_ray_diff / workspace / params must be 16-byte aligned");
  NFB_REQUIRE((long long)R * S * V < (1ll << 31), NFB_EUNSUPPORTED, "nfb_gnt_bwd: more than 2^31 rows per call; chunk the rays");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = R * S;
  const size_t rows = (size_t)N * V, NB = (size_t)N * D;
  float* F = reinterpret_cast<float*>(workspace);
  float* dF = F + rows * D;
  float* VPA = dF + rows * D;                     // depth x { VP [rows][64], A8 [rows][8] }, written by the checkpointing forward
  float* CK = VPA + (size_t)depth * rows * (D + 8);
  float* dq = CK + NB * (size_t)(5 * depth + 1);
  float* B = dq + NB;                              // B[0..6]
  auto ck = [&](int i, int j) { return CK + NB * (size_t)(5 * i + j); };
  auto buf = [&](int i) { return B + NB * (size_t)i; };
  const int sms = nfb_num_sms();
  const int out_stride = ret_alpha ? 3 + S : 3;
  int rc;

  const int ray_block = ((S + 31) / 32) * 32;
  const size_t sm_proj = (size_t)(2 * D + 4 * D * D) * sizeof(float), sm_post = (size_t)(D + 3 * D * D) * sizeof(float),
               sm_vbwd = (size_t)(VB_TOTAL + 4 * 32 * CS) * sizeof(float),
               sm_qb = (size_t)QB_TOTAL * sizeof(float), sm_eb = (size_t)EB_TOTAL * sizeof(float),
               sm_ffn = (size_t)(FS_B2 + 256 * D) * sizeof(float);     // k_gnt_ffn_bwd: weights + the per-thread dx columns
  int rpc_b = 256 / ray_block;
  if (rpc_b < 1) rpc_b = 1;
  const size_t sm_rcore = (size_t)rpc_b * (size_t)(4 * 16 + 4) * S * sizeof(float);
  if ((rc = set_smem(k_gnt_ffn_bwd, sm_ffn, "k_gnt_ffn_bwd"))) return rc;
  if ((rc = set_smem(k_gnt_proj, sm_proj, "k_gnt_proj"))) return rc;
  if ((rc = set_smem(k_gnt_post, sm_post, "k_gnt_post"))) return rc;
  if ((rc = set_smem(k_gnt_view_row_bwd, sm_vbwd, "k_gnt_view_row_bwd"))) return rc;
  if ((rc = set_smem(k_gnt_qfc_bwd, sm_qb, "k_gnt_qfc_bwd"))) return rc;
  if ((rc = set_smem(k_gnt_ray_core_bwd, sm_rcore, "k_gnt_ray_core_bwd"))) return rc;
  if ((rc = set_smem(k_gnt_embed_bwd, sm_eb, "k_gnt_embed_bwd"))) return rc;

  auto grid_n = [&](size_t n, int block, int per_sm) {
    size_t g = (n + block - 1) / block;
    if (g > (size_t)sms * per_sm) g = (size_t)sms * per_sm;
    return (int)(g < 1 ? 1 : g);
  };
  const int ffn_grid = grid_n(N, 256, 1);
  const int rcore_ctas = (R + rpc_b - 1) / rpc_b, rcore_grid = rcore_ctas < sms * 2 ? rcore_ctas : sms * 2;

  // ---------------- checkpointing forward (fp32 kernels of nfb_gnt.cu) ----------------
  if (mode != 2 && (rc = gnt_forward_checkpoints(R, S, V, depth, rgb_feat, ray_diff, mask, pts, ray_d, params, F, CK, VPA, out, ret_alpha, buf(0), st)))
    return rc;
  if (mode == 1) return NFB_OK;

  // ---------------- reverse sweep ----------------
  cudaError_t e = cudaMemsetAsync(dF, 0, rows * D * sizeof(float), st);
  if (e == cudaSuccess && d_ray_diff) e = cudaMemsetAsync(d_ray_diff, 0, rows * 4 * sizeof(float), st);
  if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "nfb_gnt_bwd: cudaMemsetAsync: %s", cudaGetErrorString(e));
  k_gnt_head_bwd<<<grid_n(N, 128, 8), 128, 0, st>>>(N, S, params + G_HEAD + (size_t)depth * L_SIZE, ck(depth, 0), d_out, out_stride, dq);
  NFB_CHECK_LAUNCH("k_gnt_head_bwd");
  for (int i = depth - 1; i >= 0; --i) {
    const float* lp = params + G_HEAD + (size_t)i * L_SIZE;
    k_gnt_ffn_bwd<<<ffn_grid, 256, sm_ffn, st>>>(N, lp + L_R_LN2_W, ck(i, 4), dq);
    NFB_CHECK_LAUNCH("k_gnt_ffn_bwd<ray>");
    const float* qd = (i & 1) == 0 ? ck(i, 3) : ck(i, 2);
    {
      ProjArgs a{};
      a.N = N; a.nw = 3; a.q_in = qd; a.dy = dq; a.ln_w = lp + L_R_LN1_W; a.ln_b = lp + L_R_LN1_B;
      a.w[0] = lp + L_R_Q; a.w[1] = lp + L_R_K; a.w[2] = lp + L_R_V; a.s0 = 0.25f; a.wo = lp + L_R_O_W;
      a.y[0] = buf(0); a.y[1] = buf(1); a.y[2] = buf(2); a.g = buf(3);
      k_gnt_proj<<<grid_n(N, 128, 3), 128, sm_proj, st>>>(a);
      NFB_CHECK_LAUNCH("k_gnt_proj<ray>");
      const float* da = (ret_alpha && i == depth - 1) ? d_out + 3 : nullptr;
      k_gnt_ray_core_bwd<<<rcore_grid, ray_block * rpc_b, sm_rcore, st>>>(R, S, rpc_b, buf(0), buf(1), buf(2), buf(3), da, out_stride,
                                                                         buf(4), buf(5), buf(6));
      NFB_CHECK_LAUNCH("k_gnt_ray_core_bwd");
      PostArgs p{};
      p.N = N; p.nw = 3; p.nv = 0; p.q_in = qd; p.dq = dq; p.ln_w = lp + L_R_LN1_W;
      p.w[0] = lp + L_R_Q; p.w[1] = lp + L_R_K; p.w[2] = lp + L_R_V; p.s0 = 0.25f;
      p.g[0] = buf(4); p.g[1] = buf(5); p.g[2] = buf(6);
      k_gnt_post<<<grid_n(N, 128, 4), 128, sm_post, st>>>(p);
      NFB_CHECK_LAUNCH("k_gnt_post<ray>");
    }
    if ((i & 1) == 0) {
      k_gnt_qfc_bwd<<<grid_n(N, 128, 2), 128, sm_qb, st>>>(N, S, pts, ray_d, lp, ck(i, 2), dq);
      NFB_CHECK_LAUNCH("k_gnt_qfc_bwd");
    }
    k_gnt_ffn_bwd<<<ffn_grid, 256, sm_ffn, st>>>(N, lp + L_V_LN2_W, ck(i, 1), dq);
    NFB_CHECK_LAUNCH("k_gnt_ffn_bwd<view>");
    {
      ProjArgs a{};
      a.N = N; a.nw = 0; a.q_in = ck(i, 0); a.dy = dq; a.ln_w = lp + L_V_LN1_W; a.ln_b = lp + L_V_LN1_B;     // only g = out_fc^T dq
      a.s0 = 1.f; a.wo = lp + L_V_O_W; a.g = buf(3);
      k_gnt_proj<<<grid_n(N, 128, 3), 128, sm_proj, st>>>(a);
      NFB_CHECK_LAUNCH("k_gnt_proj<view>");
      float* VP = VPA + (size_t)i * rows * (D + 8);           // this layer's rows, saved by the forward (consumed in place below)
      float* A8 = VP + rows * D;
      k_gnt_view_core_bwd<<<grid_n((size_t)N * 2, 128, 8), 128, 0, st>>>(N, V, A8, VP, mask, buf(3), lp);
      NFB_CHECK_LAUNCH("k_gnt_view_core_bwd");
      k_gnt_view_row_bwd<<<grid_n(rows, 128, 4), 128, sm_vbwd, st>>>(rows, A8, VP, ray_diff, lp, dF, d_ray_diff);
      NFB_CHECK_LAUNCH("k_gnt_view_row_bwd");
      PostArgs p{};
      p.N = N; p.nw = 1; p.nv = V; p.q_in = ck(i, 0); p.dq = dq; p.ln_w = lp + L_V_LN1_W; p.w[0] = lp + L_V_Q; p.s0 = 1.f; p.g[0] = VP;
      k_gnt_post<<<grid_n(N, 128, 4), 128, sm_post, st>>>(p);
      NFB_CHECK_LAUNCH("k_gnt_post<view>");
    }
  }
  k_gnt_qinit_bwd<<<grid_n(NB, 256, 16), 256, 0, st>>>(NB, V, F, dq, dF);
  NFB_CHECK_LAUNCH("k_gnt_qinit_bwd");
  k_gnt_embed_bwd<<<grid_n(rows, 128, 4), 128, sm_eb, st>>>(rows, rgb_feat, params, dF, d_rgb_feat);
  NFB_CHECK_LAUNCH("k_gnt_embed_bwd");
  return NFB_OK;
}
