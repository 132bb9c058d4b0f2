// C-ABI entry points of the IBRNet view stage (argument validation + dispatch to the instantiations).
#include "nfb_view_stage.cuh"
#include "nfb_view_tc.cuh"
#include "nfb_view_tc_bwd.cuh"
#include "nfb_view_tc_bwd2.cuh"
using nfbview::ViewArgs;

static int check_view_args(const char* who, int N, int S, int V, const float* rgb_feat, const float* ray_diff,
                    const float* mask, int H, int W, int fh, int fw, const float* xyz, const float* ray_o,
                    const float* ray_d, const float* z, const float* cam, const float* imgs, const float* feat,
                    const float* params) {
  NFB_REQUIRE(N >= 0 && V >= 1, NFB_EINVAL, "%s: bad arguments (N=%d V=%d)", who, N, V);
  NFB_REQUIRE(V <= NFB_MAX_VIEWS, NFB_EUNSUPPORTED, "%s: V=%d > %d views", who, V, NFB_MAX_VIEWS);
  if (N == 0) return NFB_OK;
  NFB_REQUIRE(params, NFB_EINVAL, "%s: params is NULL", who);
  if (rgb_feat) {
    NFB_REQUIRE(ray_diff && mask, NFB_EINVAL, "%s: tensor mode needs ray_diff and mask", who);
    NFB_REQUIRE(((uintptr_t)ray_diff % 16) == 0, NFB_EINVAL, "%s: ray_diff must be 16-byte aligned", who);
  } else {
    NFB_REQUIRE(H >= 2 && W >= 2 && fh >= 1 && fw >= 1 && cam && imgs && feat, NFB_EINVAL,
                "%s: fused mode needs image/feature sizes, cam, imgs, feat", who);
    NFB_REQUIRE(((uintptr_t)feat % 16) == 0, NFB_EINVAL, "%s: feat must be 16-byte aligned", who);
    if (!xyz)
      NFB_REQUIRE(ray_o && ray_d && z && S >= 1 && (N % S) == 0, NFB_EINVAL,
                  "%s: implicit points need ray_o, ray_d, z and S | N", who);
  }
  return NFB_OK;
}

extern "C" size_t nfb_view_stash_bytes(int N, int V) {
  if (N <= 0 || V < 1 || V > NFB_MAX_VIEWS) return 0;
  const int TS = nfbvtc::row_map(V).TS;      // the forward (SAVE) and the stash backward use the same tile mapping
  return (size_t)((N + TS - 1) / TS) * nfbvtc::ST_TILE_BYTES;
}

extern "C" int nfb_ibrnet_view_fwd(int N, int S, int V, int anti_alias, const float* rgb_feat, const float* ray_diff,
                                   const float* mask, int H, int W, int fh, int fw, const float* xyz,
                                   const float* ray_o, const float* ray_d, const float* z, const float* cam,
                                   const float* imgs, const float* feat, const float* params, float* ps,
                                   float* stash, int precision, void* stream) {
  int rc = check_view_args("nfb_ibrnet_view_fwd", N, S, V, rgb_feat, ray_diff, mask, H, W, fh, fw, xyz, ray_o, ray_d,
                           z, cam, imgs, feat, params);
  if (rc) return rc;
  if (N == 0) return NFB_OK;
  NFB_REQUIRE(ps, NFB_EINVAL, "nfb_ibrnet_view_fwd: ps is NULL");
  NFB_REQUIRE(precision >= NFB_PREC_FP32 && precision <= NFB_PREC_BF16, NFB_EINVAL, "nfb_ibrnet_view_fwd: bad precision %d", precision);
  ViewArgs a{};
  a.N = N; a.S = S; a.V = V; a.anti_alias = anti_alias;
  a.rgb_feat = rgb_feat; a.ray_diff = ray_diff; a.mask = mask;
  a.H = H; a.W = W; a.fh = fh; a.fw = fw;
  a.pts = PointSrc{xyz, ray_o, ray_d, z, S};
  a.cam = cam; a.imgs = imgs; a.feat = feat; a.params = params; a.ps = ps;
  cudaStream_t st = (cudaStream_t)stream;
  if (stash) {
    NFB_REQUIRE(!rgb_feat && precision != NFB_PREC_FP32, NFB_EUNSUPPORTED,
                "nfb_ibrnet_view_fwd: the activation stash exists for the fused tensor-core forms only");
    NFB_REQUIRE(((uintptr_t)stash % 16) == 0, NFB_EINVAL, "nfb_ibrnet_view_fwd: stash must be 16-byte aligned");
    a.stash = stash;
    return precision == NFB_PREC_BF16 ? nfb_launch_view_tc_fwd_p1_fused_save(a, st) : nfb_launch_view_tc_fwd_p3_fused_save(a, st);
  }
  if (precision == NFB_PREC_BF16X3) return rgb_feat ? nfb_launch_view_tc_fwd_p3_tensor(a, st) : nfb_launch_view_tc_fwd_p3_fused(a, st);
  if (precision == NFB_PREC_BF16) return rgb_feat ? nfb_launch_view_tc_fwd_p1_tensor(a, st) : nfb_launch_view_tc_fwd_p1_fused(a, st);
  if (rgb_feat) return nfb_launch_view_tensor_fwd(a, st);
  return nfb_launch_view_fused_fwd(a, st);
}

extern "C" int nfb_ibrnet_view_bwd(int N, int S, int V, int anti_alias, const float* rgb_feat, const float* ray_diff,
                                   const float* mask, int H, int W, int fh, int fw, const float* xyz,
                                   const float* ray_o, const float* ray_d, const float* z, const float* cam,
                                   const float* imgs, const float* feat, const float* params, const float* ps,
                                   const float* d_ps, float* d_rgb_feat, float* d_feat, float* d_imgs,
                                   const float* stash, int precision, void* stream) {
  int rc = check_view_args("nfb_ibrnet_view_bwd", N, S, V, rgb_feat, ray_diff, mask, H, W, fh, fw, xyz, ray_o, ray_d,
                           z, cam, imgs, feat, params);
  if (rc) return rc;
  if (N == 0) return NFB_OK;
  NFB_REQUIRE(ps && d_ps, NFB_EINVAL, "nfb_ibrnet_view_bwd: ps / d_ps is NULL");
  NFB_REQUIRE(precision >= NFB_PREC_FP32 && precision <= NFB_PREC_BF16, NFB_EINVAL, "nfb_ibrnet_view_bwd: bad precision %d", precision);
  if (rgb_feat) NFB_REQUIRE(d_rgb_feat, NFB_EINVAL, "nfb_ibrnet_view_bwd: tensor mode needs d_rgb_feat");
  else NFB_REQUIRE(((uintptr_t)d_feat % 16) == 0, NFB_EINVAL, "nfb_ibrnet_view_bwd: d_feat must be 16-byte aligned");
  if (N == 0) return NFB_OK;
  if (!rgb_feat && !d_feat && !d_imgs) return NFB_OK;
  ViewArgs a{};
  a.N = N; a.S = S; a.V = V; a.anti_alias = anti_alias;
  a.rgb_feat = rgb_feat; a.ray_diff = ray_diff; a.mask = mask;
  a.H = H; a.W = W; a.fh = fh; a.fw = fw;
  a.pts = PointSrc{xyz, ray_o, ray_d, z, S};
  a.cam = cam; a.imgs = imgs; a.feat = feat; a.params = params; a.ps = const_cast<float*>(ps);
  a.d_ps = d_ps; a.d_rgb_feat = d_rgb_feat; a.d_feat = d_feat; a.d_imgs = d_imgs;
  NFB_REQUIRE(((uintptr_t)ps % 16) == 0 && ((uintptr_t)d_ps % 16) == 0, NFB_EINVAL, "nfb_ibrnet_view_bwd: ps / d_ps must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (stash) {
    NFB_REQUIRE(!rgb_feat && precision != NFB_PREC_FP32, NFB_EUNSUPPORTED,
                "nfb_ibrnet_view_bwd: the activation stash exists for the fused tensor-core forms only");
    NFB_REQUIRE(((uintptr_t)stash % 16) == 0, NFB_EINVAL, "nfb_ibrnet_view_bwd: stash must be 16-byte aligned");
    a.stash = const_cast<float*>(stash);
    return precision == NFB_PREC_BF16 ? nfb_launch_view_tc_bwd_stash_p1(a, st) : nfb_launch_view_tc_bwd_stash_p3(a, st);
  }
  if (precision == NFB_PREC_BF16X3) return rgb_feat ? nfb_launch_view_tc_bwd_p3_tensor(a, st) : nfb_launch_view_tc_bwd_p3_fused(a, st);
  if (precision == NFB_PREC_BF16) return rgb_feat ? nfb_launch_view_tc_bwd_p1_tensor(a, st) : nfb_launch_view_tc_bwd_p1_fused(a, st);
  if (rgb_feat) return nfb_launch_view_tensor_bwd(a, st);
  return nfb_launch_view_fused_bwd(a, st);
}

extern "C" int nfb_ibrnet_view_wgrad(int N, int S, int V, int anti_alias, const float* rgb_feat, const float* ray_diff,
                                     const float* mask, int H, int W, int fh, int fw, const float* xyz,
                                     const float* ray_o, const float* ray_d, const float* z, const float* cam,
                                     const float* imgs, const float* feat, const float* params, const float* ps,
                                     const float* d_ps, float* d_rgb_feat, float* d_feat, float* d_imgs,
                                     float* d_params, void* stream) {
  int rc = check_view_args("nfb_ibrnet_view_wgrad", N, S, V, rgb_feat, ray_diff, mask, H, W, fh, fw, xyz, ray_o, ray_d,
                           z, cam, imgs, feat, params);
  if (rc) return rc;
  if (N == 0) return NFB_OK;
  NFB_REQUIRE(ps && d_ps && d_params, NFB_EINVAL, "nfb_ibrnet_view_wgrad: ps / d_ps / d_params is NULL");
  if (rgb_feat) NFB_REQUIRE(d_rgb_feat, NFB_EINVAL, "nfb_ibrnet_view_wgrad: tensor mode needs d_rgb_feat");
  else NFB_REQUIRE(((uintptr_t)d_feat % 16) == 0, NFB_EINVAL, "nfb_ibrnet_view_wgrad: d_feat must be 16-byte aligned");
  NFB_REQUIRE(((uintptr_t)ps % 16) == 0 && ((uintptr_t)d_ps % 16) == 0, NFB_EINVAL, "nfb_ibrnet_view_wgrad: ps / d_ps must be 16-byte aligned");
  ViewArgs a{};
  a.N = N; a.S = S; a.V = V; a.anti_alias = anti_alias;
  a.rgb_feat = rgb_feat; a.ray_diff = ray_diff; a.mask = mask;
  a.H = H; a.W = W; a.fh = fh; a.fw = fw;
  a.pts = PointSrc{xyz, ray_o, ray_d, z, S};
  a.cam = cam; a.imgs = imgs; a.feat = feat; a.params = params; a.ps = const_cast<float*>(ps);
  a.d_ps = d_ps; a.d_rgb_feat = d_rgb_feat; a.d_feat = d_feat; a.d_imgs = d_imgs; a.d_params = d_params;
  cudaStream_t st = (cudaStream_t)stream;
  return rgb_feat ? nfb_launch_view_tensor_wgrad(a, st) : nfb_launch_view_fused_wgrad(a, st);
}
