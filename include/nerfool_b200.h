/*
 * nerfool_b200 — C ABI of the B200-native per-ray generalizable-NeRF hot path.
 *
 * The reference (GATECH-EIC/NeRFool) is pure Python/PyTorch and has no FFI of its own; every entry point
 * below replaces one reference Python function on the path and cites it.  The reference-side binding a
 * maintainer would add is a ctypes stub; nerfool_b200/_lib.py is exactly that stub.
 *
 * Conventions
 *   - All pointers are DEVICE pointers (fp32 unless noted) owned by the caller (PyTorch's allocator).
 *     The library never allocates or frees device memory and keeps no mutable global state.
 *   - `stream` is a cudaStream_t passed as void*.  Calls are asynchronous and stream-ordered; no host sync.
 *   - Return value: NFB_OK (0) or a negative NfbStatus; nfb_last_error_string() gives the text
 *     (thread-local).  No exceptions cross the ABI.  There is no CPU fallback.
 *   - Row order everywhere is the reference's: points p = r*S + s (ray-major), rows (p, v) view-minor.
 *   - Feature maps are consumed CHANNEL-LAST: feat[V][fh][fw][32]; source images as the reference stores
 *     them: imgs[V][H][W][3] (sample_ray.py:124).
 *
 * Camera block `cam` (device, fp32): for each source view v, 16 floats at cam[16*v]:
 *     [0..11]  rows 0..2 of P_v = K_v * inverse(c2w_v), row-major 3x4   (projection.py:52-56)
 *     [12..14] camera centre c2w_v[:3,3]                                 (projection.py:76,80)
 *     [15]     unused
 *   followed by cam[16*V .. 16*V+2] = centre of the target (query) camera (projection.py:77-78).
 *   P_v is computed on the host with the same torch ops as the reference so that the in-frustum masks are
 *   bit-identical; the kernels evaluate P*[x,y,z,1] as the FMA chain measured to reproduce the CPU bmm.
 *
 * IBRNet parameter blob `params` (device, fp32, NFB_IBRNET_PARAM_FLOATS floats): the tensors of
 *   IBRNet.state_dict() (mlp_network.py:153-208) in torch's native [out][in] row-major layout, concatenated
 *   in the order of nfb_ibrnet_param_offset() / NFB_PARAM_* below.
 */
#ifndef NERFOOL_B200_H_
#define NERFOOL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  NFB_OK = 0,
  NFB_EINVAL = -1,        /* bad shape / null pointer / misaligned buffer */
  NFB_EUNSUPPORTED = -2,  /* shape outside what the kernels are built for (e.g. C != 35, S > 256, V > 32) */
  NFB_ECUDA = -3          /* a CUDA runtime call or launch failed; text carries cudaGetErrorString */
} NfbStatus;

/* Arithmetic of the dense layers of the IBRNet view and ray stages (`precision` argument):
 *   NFB_PREC_FP32   fp32 FMA on the CUDA cores (the exactness reference of this library)
 *   NFB_PREC_BF16X3 tcgen05 tensor cores, operands split hi+lo in bf16, 3 MMA passes, fp32 accumulation:
 *                   products exact to ~2^-17 -> results inside the reference's fp32 tolerance (default)
 *   NFB_PREC_BF16   tcgen05 tensor cores, plain bf16 operands, fp32 accumulation (throughput mode)       */
typedef enum { NFB_PREC_FP32 = 0, NFB_PREC_BF16X3 = 1, NFB_PREC_BF16 = 2 } NfbPrecision;

#define NFB_FEAT_CH 32            /* deep-feature channels per level (config.py:57-58)          */
#define NFB_ROW_CH 35             /* 3 RGB + 32 features per (sample, view) row                  */
#define NFB_PS_STRIDE 72          /* floats per sample in the view-stage -> ray-stage buffer     */
#define NFB_MAX_SAMPLES 256
#define NFB_MAX_VIEWS 32
#define NFB_IBRNET_PARAM_FLOATS 20136

/* per-sample interface buffer `ps` [N][NFB_PS_STRIDE]:
 *   [0,32) weighted mean of x over views   [32,64) weighted variance   [64] mean_v(weight)
 *   [65,68) blended RGB (rgb_out)          [68] number of valid views  [69,72) unused
 * (mlp_network.py:257-258,262,272).  The backward twin `d_ps` holds the cotangents of [0,68). */

int nfb_version(void);
const char* nfb_last_error_string(void);
/* offset (in floats) of parameter tensor `name` (e.g. "base_fc.0.weight", "s") inside the blob, or -1 */
int nfb_ibrnet_param_offset(const char* name);

/* ---- sample_along_camera_ray  (render_ray.py:73-116) -------------------------------------------------
 * z_out[R][S].  inv_uniform: uniform in 1/z.  t_rand: NULL for det=True, else [R][S] uniforms in [0,1)
 * (the host draws them with torch.rand_like so the stream matches the reference's).                    */
int nfb_coarse_depths(int R, int S, float near_depth, float far_depth, int inv_uniform,
                      const float* t_rand, float* z_out, void* stream);

/* ---- Projector.compute  (projection.py:89-132) -------------------------------------------------------
 * Points are given either explicitly (xyz[N][3], ray_o = ray_d = z = NULL, S ignored) or implicitly as
 * pts = z*ray_d + ray_o (xyz = NULL; ray_o, ray_d [R][3], z [R][S], N = R*S).
 * Outputs rgb_feat[N][V][35], ray_diff[N][V][4], mask[N][V] (0/1 floats).                               */
int nfb_project_gather_fwd(int N, int S, int V, int H, int W, int fh, int fw,
                           const float* xyz, const float* ray_o, const float* ray_d, const float* z,
                           const float* cam, const float* imgs, const float* feat,
                           float* rgb_feat, float* ray_diff, float* mask, void* stream);
/* Backward of the two bilinear gathers (grid_sampler_2d backward w.r.t. input, projection.py:119,123):
 * d_feat[V][fh][fw][32] += ..., d_imgs[V][H][W][3] += ... (either may be NULL; caller zero-fills).      */
int nfb_project_gather_bwd(int N, int S, int V, int H, int W, int fh, int fw,
                           const float* xyz, const float* ray_o, const float* ray_d, const float* z,
                           const float* cam, const float* d_rgb_feat, float* d_feat, float* d_imgs,
                           void* stream);
/* grid_sampler_2d backward w.r.t. the sampling grid: d_grid[N][V][2] = d loss / d (normalised x, y) of both gathers.
 * gnt/projection.py:84-132 keeps the source cameras in the graph (eval/gnt/eval_adv.py:749-869, --perturb_camera); the chain
 * from the grid to the camera vectors is host-side torch (nerfool_b200/ops.py: ProjectGatherCam).                          */
int nfb_project_grid_bwd(int N, int S, int V, int H, int W, int fh, int fw,
                         const float* xyz, const float* ray_o, const float* ray_d, const float* z,
                         const float* cam, const float* imgs, const float* feat, const float* d_rgb_feat,
                         float* d_grid, void* stream);

/* ---- IBRNet.forward  (mlp_network.py:222-274) --------------------------------------------------------
 * Two kernels: the view stage (per (sample,view) rows: ray_dir_fc, pooling, base_fc, vis_fc, vis_fc2,
 * rgb_fc + blending) writes ps[N][72]; the ray stage (per ray: geometry_fc, pos-enc, ray attention,
 * LayerNorm, sigma head) turns it into raw[R][S][4].
 * View-stage input is EITHER the materialised tensors (rgb_feat/ray_diff/mask as Projector.compute
 * returns them) OR, in fused mode (rgb_feat == NULL), the geometry arguments of nfb_project_gather_fwd:
 * the kernel then projects and gathers on the fly and [N][V][35] is never written.                     */
/* Activation stash (optional, fused tensor-core forms only): when `stash` is non-NULL the forward also writes what the
 * data-gradient needs per (sample, view) row -- x0, x2, a few scalars and 16-bit ELU-derivative codes, 768 B per
 * row -- and nfb_ibrnet_view_bwd given the same buffer runs as a pure backward instead of recomputing the forward
 * (on a B200 the 2 x 768 B/row of HBM traffic is several times cheaper than the recompute).  The caller allocates
 * nfb_view_stash_bytes(N, V) bytes (16-byte aligned) and keeps them until the backward has run. */
size_t nfb_view_stash_bytes(int N, int V);
int nfb_ibrnet_view_fwd(int N, int S, int V, int anti_alias,
                        const float* rgb_feat, const float* ray_diff, const float* mask,
                        int H, int W, int fh, int fw,
                        const float* xyz, const float* ray_o, const float* ray_d, const float* z,
                        const float* cam, const float* imgs, const float* feat,
                        const float* params, float* ps, float* stash, int precision, void* stream);
/* The ray stage has the same optional activation stash (tensor-core forms): 560 B per sample (q, k, v,
 * attention output and softmax statistics, LayerNorm xhat / rstd, ELU-derivative codes). */
size_t nfb_ray_stash_bytes(int R, int S);
/* pixel_mask (optional, may be NULL): uint8 [R][S] = (number of valid observations of the sample > 1), the `mask`
 * argument of raw2outputs (render_ray.py:210), written compactly so that nfb_composite_fwd does not have to pick it out of
 * the 288-byte interface rows. */
int nfb_ibrnet_ray_fwd(int R, int S, const float* ps, const float* params, const float* pos_enc /*[S][16]*/,
                       float* raw /*[R][S][4]*/, uint8_t* pixel_mask /*[R][S]*/, float* stash, int precision, void* stream);
/* Backward (data gradients): d_raw[R][S][4] -> d_ps[N][72] -> d_rgb_feat[N][V][35] (tensor mode) or a
 * scatter into d_feat / d_imgs (fused mode, rgb_feat == NULL).                                         */
int nfb_ibrnet_ray_bwd(int R, int S, const float* ps, const float* params, const float* pos_enc,
                       const float* d_raw, float* d_ps, const float* stash, int precision, void* stream);
int nfb_ibrnet_view_bwd(int N, int S, int V, int anti_alias,
                        const float* rgb_feat, const float* ray_diff, const float* mask,
                        int H, int W, int fh, int fw,
                        const float* xyz, const float* ray_o, const float* ray_d, const float* z,
                        const float* cam, const float* imgs, const float* feat,
                        const float* params, const float* ps, const float* d_ps,
                        float* d_rgb_feat, float* d_feat, float* d_imgs, const float* stash, int precision,
                        void* stream);
/* Backward WITH parameter gradients (training: train.py:317-327 back-propagates the rendering loss into the IBRNet
 * weights; mlp_network.py:153-208 lists them).  Same data-gradient outputs as the two functions above (d_feat /
 * d_imgs / d_rgb_feat may be NULL when only the weights train) and, in addition, the gradient of every tensor of
 * the parameter blob, ACCUMULATED (+=) into d_params[NFB_IBRNET_PARAM_FLOATS] in the blob's layout -- the view
 * stage fills s, ray_dir_fc, base_fc, vis_fc, vis_fc2 and rgb_fc, the ray stage geometry_fc, ray_attention.* and
 * out_geometry_fc; pos_encoding is a buffer and has no gradient.  Kernels that recompute the forward per tile in fp32
 * (no stash).  View stage: the weight gradients are GEMMs over the row index on the tensor cores (operands staged as
 * bf16 in shared memory, fp32 accumulators resident in TMEM, flushed with one float atomic per weight and CTA);
 * biases, s and the ray stage reduce per-row outer products with a warp exchange butterfly in fp32.             */
int nfb_ibrnet_ray_wgrad(int R, int S, const float* ps, const float* params, const float* pos_enc,
                         const float* d_raw, float* d_ps, float* d_params, void* stream);
int nfb_ibrnet_view_wgrad(int N, int S, int V, int anti_alias,
                          const float* rgb_feat, const float* ray_diff, const float* mask,
                          int H, int W, int fh, int fw,
                          const float* xyz, const float* ray_o, const float* ray_d, const float* z,
                          const float* cam, const float* imgs, const float* feat,
                          const float* params, const float* ps, const float* d_ps,
                          float* d_rgb_feat, float* d_feat, float* d_imgs, float* d_params, void* stream);

/* ---- GNT.forward  (gnt/transformer_network.py:270-309; SURVEY.md 8 row a15, BASELINE config 5) ------------------
 * View transformer (Attention2D: subtraction attention over the V source views, softmax per channel) + ray
 * transformer (4-head self-attention over the S samples), netwidth 64, `depth` layers, q_fc with 63-d positional
 * encodings of the points / view directions on even layers.  Inputs are what Projector.compute returns
 * (rgb_feat[R][S][V][35], ray_diff[R][S][V][4], mask[R][S][V][1]) plus pts[R][S][3] and ray_d[R][3].
 * out: [R][3] or, with ret_alpha, [R][3+S] (rgb | attention row of query 0 of the last ray transformer, the
 * "learned density" gnt/render_ray.py:249-250 turns into depth).  Forward only in this round (eval mode: dropout is
 * the identity).  precision (NfbPrecision): NFB_PREC_FP32 = CUDA-core kernels; NFB_PREC_BF16X3 / NFB_PREC_BF16 = every
 * 64-wide linear layer (q/k/v/out projections of both attentions, the feed-forward blocks) as tcgen05 128 x 64 x 64 tiles
 * with split (fp32-equivalent) or plain bf16 operands; the per-channel view softmax, the d = 16 ray attention, the
 * embedding and the positional q_fc stay on the CUDA cores.
 * params: one flat fp32 blob of nfb_gnt_param_floats(depth) floats: header (rgbfeat_fc), `depth` layer blocks
 * (view_crosstrans.i | q_fcs.i | view_selftrans.i), tail (norm, rgb_fc); nfb_gnt_param_offset(depth, name) gives the
 * offset of a header / tail tensor ("rgbfeat_fc.0.weight", "norm.bias", ...), of the first layer block ("layer0"),
 * the block size ("layer_size") and of a tensor inside a block ("view.attn.q_fc.weight", "q_fc.0.bias",
 * "ray.ff.fc2.weight", ...).  workspace: nfb_gnt_workspace_bytes(R, S, V) bytes, 16-byte aligned (projected view
 * features F, per-row k / v [R*S*V][64]; the running query q and five per-sample buffers [R*S][64]).                                                          */
int nfb_gnt_param_floats(int depth);
int nfb_gnt_param_offset(int depth, const char* name);
size_t nfb_gnt_workspace_bytes(int R, int S, int V);
int nfb_gnt_fwd(int R, int S, int V, int depth, int ret_alpha,
                const float* rgb_feat, const float* ray_diff, const float* mask, const float* pts, const float* ray_d,
                const float* params, float* out, void* workspace, size_t workspace_bytes, int precision, void* stream);

/* ---- GNT data gradient  (autograd of gnt/transformer_network.py:270-309 w.r.t. its sampled inputs; what
 * eval/gnt/eval_adv.py:282-545 back-propagates through Projector.compute to the source-image perturbation)
 * d_out: cotangent of nfb_gnt_fwd's `out` ([R][3] or [R][3+S]; the alpha columns are differentiated too).
 * d_rgb_feat[R][S][V][35] (written) and, if not NULL, d_ray_diff[R][S][V][4] (written; feeds the camera gradients of
 * gnt/projection.py:64-87, which does not detach the source cameras).  The call re-runs the forward in fp32 on the CUDA
 * cores with the running query checkpointed after every block and sweeps back block by block (ReLU masks are those of
 * the fp32 forward, i.e. the reference's to ~1e-7).  workspace: nfb_gnt_bwd_workspace_bytes(R, S, V, depth) bytes, 16-byte
 * aligned; no parameter gradients (the attack optimises the perturbation, the network is frozen: eval_adv.py:959).      */
size_t nfb_gnt_bwd_workspace_bytes(int R, int S, int V, int depth);
int nfb_gnt_bwd(int R, int S, int V, int depth, int ret_alpha,
                const float* rgb_feat, const float* ray_diff, const float* mask, const float* pts, const float* ray_d,
                const float* params, const float* d_out, float* d_rgb_feat, float* d_ray_diff,
                void* workspace, size_t workspace_bytes, void* stream);
/* The same in two calls, for callers that know at forward time that a gradient will be asked for (PyTorch autograd): nfb_gnt_fwd_save
 * is the fp32 checkpointing forward (writes `out` and fills the workspace), nfb_gnt_bwd_saved the reverse sweep on that workspace -- the
 * forward is not run twice.  Same workspace size; the workspace must stay untouched between the two calls.                          */
int nfb_gnt_fwd_save(int R, int S, int V, int depth, int ret_alpha,
                     const float* rgb_feat, const float* ray_diff, const float* mask, const float* pts, const float* ray_d,
                     const float* params, float* out, void* workspace, size_t workspace_bytes, void* stream);
int nfb_gnt_bwd_saved(int R, int S, int V, int depth, int ret_alpha,
                      const float* rgb_feat, const float* ray_diff, const float* mask, const float* pts, const float* ray_d,
                      const float* params, const float* d_out, float* d_rgb_feat, float* d_ray_diff,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- raw2outputs  (render_ray.py:123-170) ------------------------------------------------------------
 * pixel_mask: uint8 [R][S] (the `mask` argument), or NULL with n_valid (stride n_valid_stride floats per
 * sample) from which pixel_mask = n_valid > 1 (render_ray.py:210).  Outputs rgb[R][3], depth[R],
 * weights[R][S], alpha[R][S], ray_mask uint8 [R].                                                      */
int nfb_composite_fwd(int R, int S, int white_bkgd, const float* raw, const float* z,
                      const uint8_t* pixel_mask, const float* n_valid, int n_valid_stride,
                      float* rgb, float* depth, float* weights, float* alpha, uint8_t* ray_mask,
                      void* stream);
/* any of d_rgb[R][3], d_depth[R], d_weights[R][S], d_alpha[R][S] may be NULL -> d_raw[R][S][4] */
int nfb_composite_bwd(int R, int S, int white_bkgd, const float* raw, const float* z,
                      const float* d_rgb, const float* d_depth, const float* d_weights,
                      const float* d_alpha, float* d_raw, void* stream);

/* ---- sample_pdf  (render_ray.py:24-70) ---------------------------------------------------------------
 * bins[R][M+1], weights[R][M] (NOT modified; the +1e-5 is applied internally), u[u_rows][n] with
 * u_rows in {1, R}.  Outputs samples[R][n] and (optional) above[R][n] int64 bin indices.               */
int nfb_sample_pdf(int R, int M, int n, const float* bins, const float* weights, const float* u,
                   int u_rows, float* samples, int64_t* above, void* stream);
/* ---- fine depths: the whole of render_ray.py:216-238 -------------------------------------------------
 * z_coarse[R][S], weights_coarse[R][S] -> z_fine[R][S+n_imp] sorted ascending.                         */
int nfb_fine_depths(int R, int S, int n_imp, int inv_uniform, const float* z_coarse,
                    const float* weights_coarse, const float* u, int u_rows, float* z_fine, void* stream);

/* ---- forward_warp  (eval/ibrnet/eval_adv.py:97-197, the Python z-buffer loop of the depth / camera consistency losses) ----
 * x_res, y_res: int32 [H*W] destination pixel of every source pixel (clamped and truncated as eval_adv.py:131-136 does);
 * depth_src float [H*W], rgb_ref float [H*W][3]; allowed: uint8 [H*W] destination pixels that accept writes (the
 * `if inds in selected_inds` test of the src2tar branch, :142) or NULL; sources: int32 [n_src] source pixels in loop order (the
 * tar2src branch iterates selected_inds, :160) or NULL = all pixels in raster order.  Outputs new_rgb [H*W][3], new_depth [H*W]
 * (initialised by the call); keys: uint64 [H*W] workspace; flag: int32, set to 1 when a depth <= 0 / NaN was met, in which case
 * the outputs are NOT valid and the caller repeats the call with sequential = 1 (the reference loop in one thread).          */
int nfb_forward_warp(int H, int W, const int* x_res, const int* y_res, const float* depth_src, const float* rgb_ref,
                     const uint8_t* allowed, const int* sources, int n_src, float* new_rgb, float* new_depth,
                     unsigned long long* keys, int* flag, int sequential, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERFOOL_B200_H_ */
