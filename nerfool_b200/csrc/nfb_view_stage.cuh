// IBRNet view stage (fp32 CUDA-core form): everything of IBRNet.forward that lives on (sample, view) rows
// -- ray_dir_fc, anti-alias pooling weights, mean/variance pooling, base_fc, vis_fc, vis_fc2, the second
// pooling, rgb_fc and the colour blending (mlp_network.py:231-258, 267-272) -- and its data-gradient.
//
// Mapping: one thread per (sample, view) row; 128-thread row groups own tiles of TS = 128/V samples
// (rows of a sample are consecutive threads); cross-view reductions go through a per-group exchange
// buffer in shared memory guarded by named barriers; all layer weights are resident in shared memory,
// transposed (nfb_dense.cuh).  Input rows come either from the materialised Projector.compute tensors or,
// in fused mode, straight from projection + bilinear gather (nfb_geom.cuh) so [N][V][35] never exists.
#pragma once
#include "nfb_dense.cuh"
#include "nfb_geom.cuh"
#include "nfb_wgrad_tc.cuh"

namespace nfbview {

constexpr int GROUP = 128;   // threads (= rows) per row group
constexpr int EXS = 37;      // exchange-buffer row stride in floats (odd => conflict-free row access)
constexpr int TS_MAX = 64;   // samples per tile (cap; reached for V <= 2)
constexpr int MVS = 72;      // per-sample pooled-statistics stride: mean0[35] at 0, var0[35] at 36

// shared-memory weight layout (floats); every offset is a multiple of 4 (LDS.128 alignment)
enum : int {
  W_DIR0 = 0,                    // [4][16]
  B_DIR0 = W_DIR0 + 4 * 16,      // 16
  W_DIR2 = B_DIR0 + 16,          // [16][36]
  B_DIR2 = W_DIR2 + 16 * 36,     // 36
  W_BASE0 = B_DIR2 + 36,         // [105][64]
  B_BASE0 = W_BASE0 + 105 * 64,  // 64
  W_BASE2 = B_BASE0 + 64,        // [64][32]
  B_BASE2 = W_BASE2 + 64 * 32,   // 32
  W_VIS0 = B_BASE2 + 32,         // [32][32]
  B_VIS0 = W_VIS0 + 32 * 32,     // 32
  W_VIS2 = B_VIS0 + 32,          // [32][36]
  B_VIS2 = W_VIS2 + 32 * 36,     // 36
  W_VISB0 = B_VIS2 + 36,         // [32][32]
  B_VISB0 = W_VISB0 + 32 * 32,   // 32
  W_VISB2 = B_VISB0 + 32,        // [32]   (single output row)
  B_VISB2 = W_VISB2 + 32,        // 4
  W_RGB0 = B_VISB2 + 4,          // [37][16]
  B_RGB0 = W_RGB0 + 37 * 16,     // 16
  W_RGB2 = B_RGB0 + 16,          // [16][8]
  B_RGB2 = W_RGB2 + 16 * 8,      // 8
  W_RGB4 = B_RGB2 + 8,           // [8]
  B_RGB4 = W_RGB4 + 8,           // 4
  W_S = B_RGB4 + 4,              // 4 (|s| in slot 0)
  W_TOTAL = W_S + 4
};

static __device__ void load_view_weights(float* sw, const float* __restrict__ p, int tid, int nt) {
  load_wt_transposed(sw + W_DIR0, p + P_DIR0_W, 16, 4, 16, tid, nt);
  load_vec_padded(sw + B_DIR0, p + P_DIR0_B, 16, 16, tid, nt);
  load_wt_transposed(sw + W_DIR2, p + P_DIR2_W, 35, 16, 36, tid, nt);
  load_vec_padded(sw + B_DIR2, p + P_DIR2_B, 35, 36, tid, nt);
  load_wt_transposed(sw + W_BASE0, p + P_BASE0_W, 64, 105, 64, tid, nt);
  load_vec_padded(sw + B_BASE0, p + P_BASE0_B, 64, 64, tid, nt);
  load_wt_transposed(sw + W_BASE2, p + P_BASE2_W, 32, 64, 32, tid, nt);
  load_vec_padded(sw + B_BASE2, p + P_BASE2_B, 32, 32, tid, nt);
  load_wt_transposed(sw + W_VIS0, p + P_VIS0_W, 32, 32, 32, tid, nt);
  load_vec_padded(sw + B_VIS0, p + P_VIS0_B, 32, 32, tid, nt);
  load_wt_transposed(sw + W_VIS2, p + P_VIS2_W, 33, 32, 36, tid, nt);
  load_vec_padded(sw + B_VIS2, p + P_VIS2_B, 33, 36, tid, nt);
  load_wt_transposed(sw + W_VISB0, p + P_VISB0_W, 32, 32, 32, tid, nt);
  load_vec_padded(sw + B_VISB0, p + P_VISB0_B, 32, 32, tid, nt);
  load_vec_padded(sw + W_VISB2, p + P_VISB2_W, 32, 32, tid, nt);
  load_vec_padded(sw + B_VISB2, p + P_VISB2_B, 1, 4, tid, nt);
  load_wt_transposed(sw + W_RGB0, p + P_RGB0_W, 16, 37, 16, tid, nt);
  load_vec_padded(sw + B_RGB0, p + P_RGB0_B, 16, 16, tid, nt);
  load_wt_transposed(sw + W_RGB2, p + P_RGB2_W, 8, 16, 8, tid, nt);
  load_vec_padded(sw + B_RGB2, p + P_RGB2_B, 8, 8, tid, nt);
  load_vec_padded(sw + W_RGB4, p + P_RGB4_W, 8, 8, tid, nt);
  load_vec_padded(sw + B_RGB4, p + P_RGB4_B, 1, 4, tid, nt);
  if (tid == 0) sw[W_S] = fabsf(__ldg(p + P_S));
}

// WG (training) kernels: TMEM accumulator columns of the layers whose parameter gradient runs on the tensor cores
// (nfb_wgrad_tc.cuh); the two single-output layers and s keep a small shared-memory accumulator (warp butterfly).
enum : int { WC_DIR0 = 0, WC_DIR2 = 16, WC_BASE0 = 64, WC_BASE2 = 128, WC_VIS0 = 160, WC_VIS2 = 192, WC_VISB0 = 240,
             WC_RGB0 = 272, WC_RGB2 = 288, WC_TOTAL = 304 };
enum : int { SG_VISB2 = 0, SG_VISB2_B = 32, SG_RGB4 = 36, SG_RGB4_B = 44, SG_S = 48,
             // bias gradients (plain row sums) stay exact fp32: warp butterfly into these slots
             SG_DIR0_B = 52, SG_DIR2_B = SG_DIR0_B + 16, SG_BASE0_B = SG_DIR2_B + 36, SG_BASE2_B = SG_BASE0_B + 64,
             SG_VIS0_B = SG_BASE2_B + 32, SG_VIS2_B = SG_VIS0_B + 32, SG_VISB0_B = SG_VIS2_B + 36, SG_RGB0_B = SG_VISB0_B + 32,
             SG_RGB2_B = SG_RGB0_B + 16, SG_TOTAL = SG_RGB2_B + 8 };
static_assert(SG_TOTAL % 4 == 0, "staging tiles behind the accumulators need 16-byte alignment");

struct ViewArgs {
  int N, S, V, anti_alias;
  // tensor mode
  const float* rgb_feat; const float* ray_diff; const float* mask;
  // fused mode
  int H, W, fh, fw;
  PointSrc pts;
  const float* cam; const float* imgs; const float* feat;
  // common
  const float* params;
  float* ps;                 // forward: output.  backward: forward output of the same inputs (read)
  // backward only
  const float* d_ps;
  float* d_rgb_feat; float* d_feat; float* d_imgs;
  // tensor-core fused mode: activation stash written by the forward, read by the backward (nfb_view_tc.cuh)
  float* stash;
  // training: parameter-gradient blob (NFB_IBRNET_PARAM_FLOATS, torch layout), accumulated into (WG kernels only)
  float* d_params;
};

// ---------------------------------------------------------------------------------------------------
// one kernel body for forward (BWD=false) and forward-recompute + backward (BWD=true)
// ---------------------------------------------------------------------------------------------------
template <bool FUSED, bool BWD, int NG, bool WG = false>
__global__ void __launch_bounds__(GROUP * NG, 1) k_view_stage(ViewArgs a) {
  static_assert(!WG || BWD, "parameter gradients are part of the backward");
  extern __shared__ __align__(16) float smem[];
  float* sw = smem;
  float* s_cam = smem + W_TOTAL;                               // 16*V+4 floats (fused mode)
  float* ex_all = s_cam + (16 * NFB_MAX_VIEWS + 4);
  const int grp = threadIdx.x / GROUP, tg = threadIdx.x % GROUP;
  float* ex = ex_all + (size_t)grp * GROUP * EXS;
  float* mv = ex_all + (size_t)NG * GROUP * EXS + (size_t)grp * TS_MAX * MVS;   // per-sample mean0 / var0
  const int bar_id = 1 + grp;
  // WG: small shared accumulators (single-output layers, s), then the bf16 staging tiles of the tensor-core
  // parameter-gradient GEMMs (one A and one B tile per group), mbarriers and the TMEM slot
  float* sg = ex_all + (size_t)NG * GROUP * EXS + (size_t)NG * TS_MAX * MVS;
  uint8_t* s_tiles = reinterpret_cast<uint8_t*>(sg + SG_TOTAL);
  uint64_t* s_wbar = reinterpret_cast<uint64_t*>(s_tiles + (size_t)NG * (nfbwg::A_TILE_BYTES + nfbwg::B_TILE_BYTES));
  uint32_t* s_wtmem = reinterpret_cast<uint32_t*>(s_wbar + NG);
  float ds_acc = 0.f;                                          // d loss / d |s| of this thread's rows
  nfbwg::WgTc wg{};
  if (WG) {
    for (int i = threadIdx.x; i < SG_TOTAL; i += blockDim.x) sg[i] = 0.f;
    // stale bf16 patterns in never-written operand rows only reach unused accumulator lanes, but they must be finite
    for (int i = threadIdx.x; i < NG * (nfbwg::A_TILE_BYTES + nfbwg::B_TILE_BYTES) / 4; i += blockDim.x)
      reinterpret_cast<uint32_t*>(s_tiles)[i] = 0u;
    if (threadIdx.x < 32) nfbtc::tmem_alloc(s_wtmem, 512);
    if (threadIdx.x == 0) {
      for (int g = 0; g < NG; ++g) nfbtc::mbar_init(s_wbar + g, 1);
      nfbtc::mbar_init_fence();
    }
  }

  load_view_weights(sw, a.params, threadIdx.x, blockDim.x);
  if (FUSED)
    for (int i = threadIdx.x; i < 16 * a.V + 3; i += blockDim.x) s_cam[i] = __ldg(a.cam + i);
  if (WG) nfbtc::fence_before_sync();
  __syncthreads();
  if (WG) {
    nfbtc::fence_after_sync();
    wg.sA = s_tiles + (size_t)grp * nfbwg::A_TILE_BYTES;
    wg.sB = s_tiles + (size_t)NG * nfbwg::A_TILE_BYTES + (size_t)grp * nfbwg::B_TILE_BYTES;
    wg.mbar = s_wbar + grp;
    wg.tmem = *s_wtmem;
    wg.phase = 0;
    wg.pending = false;
    wg.tg = tg;
    wg.bar_id = bar_id;
    if (threadIdx.x < 128) nfbwg::wg_zero(wg.tmem, threadIdx.x >> 5, WC_TOTAL);     // warps 0..3 own TMEM lanes 0..127
    nfbtc::fence_before_sync();
    __syncthreads();
    nfbtc::fence_after_sync();
  }

  const int V = a.V;
  const int TS = (GROUP / V < TS_MAX) ? GROUP / V : TS_MAX;    // samples per tile
  const int sl = tg / V, v = tg - sl * V;
  const int ntiles = (a.N + TS - 1) / TS;
  const float Wm1 = (float)a.W - 1.f, Hm1 = (float)a.H - 1.f;
  const float s_abs = sw[W_S];

  for (int tile = blockIdx.x * NG + grp; tile < ntiles; tile += gridDim.x * NG) {
    const int p = tile * TS + sl;                              // global sample (point) index
    const bool active = (sl < TS) && (p < a.N);
    const int base = active ? sl * V : 0;                      // first exchange row of this sample
    float* mvs = mv + (active ? sl : 0) * MVS;                 // pooled statistics of this sample
    const size_t row = (size_t)p * V + v;                      // global (sample, view) row

    // ---------------- input row: x[35], rd[4], mask ----------------
    float x[NFB_ROW_CH];
    float rd[4];
    float mk = 0.f;
    float gx = 0.f, gy = 0.f;
    if (FUSED) {
      if (active) {
        float X, Y, Z;
        load_point(a.pts, p, X, Y, Z);
        const ViewGeom g = view_geometry(X, Y, Z, s_cam + 16 * v, s_cam + 16 * V, Wm1, Hm1);
        gather_row(g, v, a.H, a.W, a.fh, a.fw, a.imgs, a.feat, x);
        rd[0] = g.rd[0]; rd[1] = g.rd[1]; rd[2] = g.rd[2]; rd[3] = g.rd[3];
        mk = g.mask; gx = g.gx; gy = g.gy;
      } else {
#pragma unroll
        for (int c = 0; c < NFB_ROW_CH; ++c) x[c] = 0.f;
        rd[0] = rd[1] = rd[2] = rd[3] = 0.f;
      }
    } else {
      // the tile's rows are contiguous in rgb_feat: coalesced copy through the exchange buffer
      const size_t row0 = (size_t)tile * TS * V;
      const size_t total = (size_t)a.N * V;
      const int rows_here = (int)((total - row0 < (size_t)(TS * V)) ? (total - row0) : (size_t)(TS * V));
      const float* src = a.rgb_feat + row0 * NFB_ROW_CH;
      for (int i = tg; i < rows_here * NFB_ROW_CH; i += GROUP) {
        const int rr = i / NFB_ROW_CH, cc = i - rr * NFB_ROW_CH;
        ex[rr * EXS + cc] = __ldg(src + i);
      }
      named_bar_sync(bar_id, GROUP);
#pragma unroll
      for (int c = 0; c < NFB_ROW_CH; ++c) x[c] = active ? ex[tg * EXS + c] : 0.f;
      if (active) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(a.ray_diff) + row);
        rd[0] = q.x; rd[1] = q.y; rd[2] = q.z; rd[3] = q.w;
        mk = __ldg(a.mask + row);
      } else {
        rd[0] = rd[1] = rd[2] = rd[3] = 0.f;
      }
      named_bar_sync(bar_id, GROUP);
    }
    const float rgb_in0 = x[0], rgb_in1 = x[1], rgb_in2 = x[2];

    // ---------------- ray_dir_fc (mlp_network.py:231) and x0 = rgb_feat + direction_feat (:233) -------
    {
      float a1[16];
      load_bias<16>(a1, sw + B_DIR0);
      dense_acc<4, 16>(sw + W_DIR0, rd, a1);
      elu_inplace<16>(a1);
      float df[36];
      load_bias<36>(df, sw + B_DIR2);
      dense_acc<16, 36>(sw + W_DIR2, a1, df);
#pragma unroll
      for (int c = 0; c < NFB_ROW_CH; ++c) x[c] += elu_f(df[c]);
    }

    // ---------------- pooling weights (:234-241) ----------------
    float w, n_valid;
    float e_aa = 1.f, mn_aa = 0.f, den_aa = 1.f;               // WG: exp term, its minimum over views, normaliser
    {
      // exp_dot_prod = exp(|s| * (dot - 1)) (:236).  The weights are DIFFERENCES of these exponentials
      // (:237), which cancel when the source views see the point under similar angles, so the exponential
      // is evaluated in fp64 and rounded once (correctly rounded, = torch's CPU result on ~99 % of inputs).
      const float e = a.anti_alias ? (float)exp((double)__fmul_rn(s_abs, __fsub_rn(rd[3], 1.f))) : 1.f;
      ex[tg * EXS + 35] = e;
      ex[tg * EXS + 36] = mk;
      named_bar_sync(bar_id, GROUP);
      float mn = 3.4e38f, nv = 0.f;
      for (int u = 0; u < V; ++u) {
        mn = fminf(mn, ex[(base + u) * EXS + 35]);
        nv += ex[(base + u) * EXS + 36];
      }
      if (!a.anti_alias) mn = 0.f;
      float sum = 0.f;
      for (int u = 0; u < V; ++u) sum += (ex[(base + u) * EXS + 35] - mn) * ex[(base + u) * EXS + 36];
      w = (e - mn) * mk / (sum + 1e-8f);
      n_valid = nv;
      e_aa = e; mn_aa = mn; den_aa = sum + 1e-8f;
      named_bar_sync(bar_id, GROUP);
    }

    // ---------------- exchange x0 / w, then base_fc with on-the-fly mean / variance (:244-248) --------
#pragma unroll
    for (int c = 0; c < NFB_ROW_CH; ++c) ex[tg * EXS + c] = x[c];
    ex[tg * EXS + 35] = w;
    named_bar_sync(bar_id, GROUP);

    // pooled mean / variance (fused_mean_variance, :145-149): channels split across the V row threads
    float wsum0 = 0.f;
    for (int u = 0; u < V; ++u) wsum0 += ex[(base + u) * EXS + 35];
    if (active) {
      for (int c = v; c < NFB_ROW_CH; c += V) {
        float m = 0.f;
        for (int u = 0; u < V; ++u) m = fmaf(ex[(base + u) * EXS + c], ex[(base + u) * EXS + 35], m);
        float vr = 0.f;
        for (int u = 0; u < V; ++u) {
          const float d = ex[(base + u) * EXS + c] - m;
          vr = fmaf(ex[(base + u) * EXS + 35] * d, d, vr);
        }
        mvs[c] = m;
        mvs[36 + c] = vr;
      }
    }
    named_bar_sync(bar_id, GROUP);

    float h1[64];
    load_bias<64>(h1, sw + B_BASE0);
#pragma unroll
    for (int c = 0; c < NFB_ROW_CH; ++c) {
      axpy_row<64>(h1, mvs[c], sw + W_BASE0 + c * 64);
      axpy_row<64>(h1, mvs[36 + c], sw + W_BASE0 + (35 + c) * 64);
      axpy_row<64>(h1, x[c], sw + W_BASE0 + (70 + c) * 64);
    }
    elu_inplace<64>(h1);
    float x1[32];
    load_bias<32>(x1, sw + B_BASE2);
    dense_acc<64, 32>(sw + W_BASE2, h1, x1);
    elu_inplace<32>(x1);

    // ---------------- vis_fc (:250-253) ----------------
    float hv[32], xv[36];
    {
      float t[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) t[c] = x1[c] * w;
      load_bias<32>(hv, sw + B_VIS0);
      dense_acc<32, 32>(sw + W_VIS0, t, hv);
      elu_inplace<32>(hv);
      load_bias<36>(xv, sw + B_VIS2);
      dense_acc<32, 36>(sw + W_VIS2, hv, xv);
      elu_inplace<36>(xv);
    }
    const float sg1 = sigmoid_f(xv[32]);
    const float vis1 = sg1 * mk;
    float x2[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) x2[c] = x1[c] + xv[c];

    // ---------------- vis_fc2 (:254) ----------------
    float hv2[32];
    float sg2, vis2;
    {
      float t[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) t[c] = x2[c] * vis1;
      load_bias<32>(hv2, sw + B_VISB0);
      dense_acc<32, 32>(sw + W_VISB0, t, hv2);
      elu_inplace<32>(hv2);
      sg2 = sigmoid_f(dot_row<32>(hv2, sw + W_VISB2) + sw[B_VISB2]);
      vis2 = sg2 * mk;
    }

    // ---------------- rgb_fc on [x2, vis2, ray_diff] (:268-270) ----------------
    float g1[16], g2[8];
    float logit;
    {
      load_bias<16>(g1, sw + B_RGB0);
      dense_acc<32, 16>(sw + W_RGB0, x2, g1);
      axpy_row<16>(g1, vis2, sw + W_RGB0 + 32 * 16);
      dense_acc<4, 16>(sw + W_RGB0 + 33 * 16, rd, g1);
      elu_inplace<16>(g1);
      load_bias<8>(g2, sw + B_RGB2);
      dense_acc<16, 8>(sw + W_RGB2, g1, g2);
      elu_inplace<8>(g2);
      logit = dot_row<8>(g2, sw + W_RGB4) + sw[B_RGB4];
      if (mk == 0.f) logit = -1e9f;
    }

    // ---------------- second exchange: x2, vis2, logit, rgb_in ----------------
    named_bar_sync(bar_id, GROUP);              // everyone is done reading x0 / w from the buffer
    if (!BWD) {
#pragma unroll
      for (int c = 0; c < 32; ++c) ex[tg * EXS + c] = x2[c];
    }
    ex[tg * EXS + 32] = vis2;
    ex[tg * EXS + 33] = logit;
    ex[tg * EXS + 34] = rgb_in0;
    ex[tg * EXS + 35] = rgb_in1;
    ex[tg * EXS + 36] = rgb_in2;
    named_bar_sync(bar_id, GROUP);

    float D = 1e-8f;
    for (int u = 0; u < V; ++u) D += ex[(base + u) * EXS + 32];
    // softmax over the views of this sample (:271)
    float mx = -3.4e38f;
    for (int u = 0; u < V; ++u) mx = fmaxf(mx, ex[(base + u) * EXS + 33]);
    float se = 0.f;
    for (int u = 0; u < V; ++u) se += __expf(ex[(base + u) * EXS + 33] - mx);
    const float inv_se = 1.f / se;

    if (!BWD) {
      // weighted mean / variance of x2 over views (:256-258), channels split across the V threads
      if (active) {
        float* out = a.ps + (size_t)p * NFB_PS_STRIDE;
        for (int c = v; c < 32; c += V) {
          float m = 0.f;
          for (int u = 0; u < V; ++u) m = fmaf(ex[(base + u) * EXS + c], ex[(base + u) * EXS + 32] / D, m);
          float vr = 0.f;
          for (int u = 0; u < V; ++u) {
            const float d = ex[(base + u) * EXS + c] - m;
            vr = fmaf((ex[(base + u) * EXS + 32] / D) * d, d, vr);
          }
          out[PS_MEAN + c] = m;
          out[PS_VAR + c] = vr;
        }
        if (v == 0) {
          float wsum = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f;
          for (int u = 0; u < V; ++u) {
            wsum += ex[(base + u) * EXS + 32] / D;
            const float b = __expf(ex[(base + u) * EXS + 33] - mx) * inv_se;
            r0 = fmaf(b, ex[(base + u) * EXS + 34], r0);
            r1 = fmaf(b, ex[(base + u) * EXS + 35], r1);
            r2 = fmaf(b, ex[(base + u) * EXS + 36], r2);
          }
          out[PS_WMEAN] = wsum / (float)V;
          out[PS_RGB + 0] = r0; out[PS_RGB + 1] = r1; out[PS_RGB + 2] = r2;
          out[PS_NVALID] = n_valid;
          out[69] = 0.f; out[70] = 0.f; out[71] = 0.f;
        }
      }
      named_bar_sync(bar_id, GROUP);            // buffer is reused by the next tile
      continue;
    }

    // =================================== backward ===================================
    if (BWD) {
      const float gate = active ? 1.f : 0.f;       // WG: rows outside the problem contribute nothing
      float d_w = 0.f;                             // WG: d loss / d (first pooling weight of this row)
      const float* dps = a.d_ps + (size_t)(active ? p : 0) * NFB_PS_STRIDE;
      const float* fps = a.ps + (size_t)(active ? p : 0) * NFB_PS_STRIDE;
      const float w2 = vis2 / D;
      float w2sum = 0.f;
      for (int u = 0; u < V; ++u) w2sum += ex[(base + u) * EXS + 32] / D;
      const float d_r0 = __ldg(dps + PS_RGB + 0), d_r1 = __ldg(dps + PS_RGB + 1), d_r2 = __ldg(dps + PS_RGB + 2);
      const float d_wmean = __ldg(dps + PS_WMEAN);

      // (1) blending softmax: d logit_v = blend_v (t_v - sum_u blend_u t_u), zero where masked_fill hit
      const float blend = __expf(logit - mx) * inv_se;
      float bt = 0.f;
      for (int u = 0; u < V; ++u) {
        const float b = __expf(ex[(base + u) * EXS + 33] - mx) * inv_se;
        const float tu = ex[(base + u) * EXS + 34] * d_r0 + ex[(base + u) * EXS + 35] * d_r1 + ex[(base + u) * EXS + 36] * d_r2;
        bt = fmaf(b, tu, bt);
      }
      const float tv = rgb_in0 * d_r0 + rgb_in1 * d_r1 + rgb_in2 * d_r2;
      const float d_logit = (mk != 0.f) ? blend * (tv - bt) : 0.f;

      // (2) rgb_fc backward -> d_x2, d_vis2
      float d_x2[32];
      float d_vis2;
      {
        float dg2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dg2[j] = d_logit * sw[W_RGB4 + j] * elu_grad_from_out(g2[j]);
        float dg1[16];
        dense_T<16, 8>(sw + W_RGB2, dg2, dg1);
#pragma unroll
        for (int k = 0; k < 16; ++k) dg1[k] *= elu_grad_from_out(g1[k]);
        dense_T<32, 16>(sw + W_RGB0, dg1, d_x2);
        d_vis2 = dot_row<16>(dg1, sw + W_RGB0 + 32 * 16);
        if (WG) {
          wgrad_rowsum<8>(sg + SG_RGB4, g2, d_logit * gate);
          const float db4[4] = {d_logit, 0.f, 0.f, 0.f};
          wgrad_rowsum<4>(sg + SG_RGB4_B, db4, gate);
          nfbwg::wgrad_tc<16, 8>(wg, WC_RGB2, g1, dg2, gate);
          wgrad_rowsum<8>(sg + SG_RGB2_B, dg2, gate);
          float xin[37];
#pragma unroll
          for (int c = 0; c < 32; ++c) xin[c] = x2[c];
          xin[32] = vis2; xin[33] = rd[0]; xin[34] = rd[1]; xin[35] = rd[2]; xin[36] = rd[3];
          nfbwg::wgrad_tc<37, 16>(wg, WC_RGB0, xin, dg1, gate);
          wgrad_rowsum<16>(sg + SG_RGB0_B, dg1, gate);
        }
      }

      // (3) second pooling backward (fused_mean_variance + weight normalisation, :255-258)
      {
        float d_w2 = d_wmean / (float)V;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float dm = __ldg(dps + PS_MEAN + c), dv = __ldg(dps + PS_VAR + c);
          const float mean = __ldg(fps + PS_MEAN + c);
          const float diff = x2[c] - mean;
          d_w2 = fmaf(dm, x2[c], d_w2);
          d_w2 = fmaf(dv * diff, diff, d_w2);
          d_x2[c] += w2 * (dm - 2.f * dv * mean * (1.f - w2sum)) + 2.f * w2 * diff * dv;
        }
        named_bar_sync(bar_id, GROUP);          // all reads of slots 33..36 above are done
        ex[tg * EXS + 33] = d_w2 * vis2;
        named_bar_sync(bar_id, GROUP);
        float sdv = 0.f;
        for (int u = 0; u < V; ++u) sdv += ex[(base + u) * EXS + 33];
        d_vis2 += d_w2 / D - sdv / (D * D);
      }

      // (4) vis_fc2 backward
      float d_vis1;
      {
        const float dz = d_vis2 * mk * sg2 * (1.f - sg2);
        float dh[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) dh[k] = dz * sw[W_VISB2 + k] * elu_grad_from_out(hv2[k]);
        float dt[32];
        dense_T<32, 32>(sw + W_VISB0, dh, dt);
        if (WG) {
          wgrad_rowsum<32>(sg + SG_VISB2, hv2, dz * gate);
          const float dbz[4] = {dz, 0.f, 0.f, 0.f};
          wgrad_rowsum<4>(sg + SG_VISB2_B, dbz, gate);
          float t2[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) t2[c] = x2[c] * vis1;
          nfbwg::wgrad_tc<32, 32>(wg, WC_VISB0, t2, dh, gate);
          wgrad_rowsum<32>(sg + SG_VISB0_B, dh, gate);
        }
        d_vis1 = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          d_vis1 = fmaf(dt[c], x2[c], d_vis1);
          d_x2[c] = fmaf(dt[c], vis1, d_x2[c]);
        }
      }

      // (5) x2 = x1 + x_res, vis1 = sigmoid(xv[32]) * mask ; vis_fc backward
      float d_x1[32];
      {
        float dxv[36];
#pragma unroll
        for (int c = 0; c < 32; ++c) dxv[c] = d_x2[c] * elu_grad_from_out(xv[c]);
        dxv[32] = d_vis1 * mk * sg1 * (1.f - sg1) * elu_grad_from_out(xv[32]);
        dxv[33] = dxv[34] = dxv[35] = 0.f;
        float dh[32];
        dense_T<32, 36>(sw + W_VIS2, dxv, dh);
#pragma unroll
        for (int k = 0; k < 32; ++k) dh[k] *= elu_grad_from_out(hv[k]);
        float dt[32];
        dense_T<32, 32>(sw + W_VIS0, dh, dt);
#pragma unroll
        for (int c = 0; c < 32; ++c) d_x1[c] = fmaf(dt[c], w, d_x2[c]);
        if (WG) {
          nfbwg::wgrad_tc<32, 36>(wg, WC_VIS2, hv, dxv, gate);
          wgrad_rowsum<36>(sg + SG_VIS2_B, dxv, gate);
          float t1[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            t1[c] = x1[c] * w;
            d_w = fmaf(dt[c], x1[c], d_w);         // vis_fc sees x * weight (:250)
          }
          nfbwg::wgrad_tc<32, 32>(wg, WC_VIS0, t1, dh, gate);
          wgrad_rowsum<32>(sg + SG_VIS0_B, dh, gate);
        }
      }

      // (6) base_fc backward -> d[mean0 | var0 | x0]
      float d_h1[64];
      {
#pragma unroll
        for (int c = 0; c < 32; ++c) d_x1[c] *= elu_grad_from_out(x1[c]);
        dense_T<64, 32>(sw + W_BASE2, d_x1, d_h1);
#pragma unroll
        for (int k = 0; k < 64; ++k) d_h1[k] *= elu_grad_from_out(h1[k]);
      }
      float m0[NFB_ROW_CH];                        // WG: mean0 of this sample (the buffer is overwritten in (7))
      if (WG) {
        nfbwg::wgrad_tc<64, 32>(wg, WC_BASE2, h1, d_x1, gate);
        wgrad_rowsum<32>(sg + SG_BASE2_B, d_x1, gate);
        float xin[105];
#pragma unroll
        for (int c = 0; c < NFB_ROW_CH; ++c) {
          m0[c] = mvs[c];
          xin[c] = mvs[c]; xin[35 + c] = mvs[36 + c]; xin[70 + c] = x[c];
        }
        nfbwg::wgrad_tc<105, 64>(wg, WC_BASE0, xin, d_h1, gate);
        wgrad_rowsum<64>(sg + SG_BASE0_B, d_h1, gate);
      }

      // (7) first pooling backward.  With Dm_c = sum_v d mean0_vc, Dv_c = sum_v d var0_vc:
      //   d x0_vc = dx_vc + w_v (Dm_c - 2 Dv_c mean0_c (1 - wsum)) + 2 w_v (x0_vc - mean0_c) Dv_c
      //           = dx_vc + w_v A_c + w_v x0_vc B_c,  B_c = 2 Dv_c,  A_c = Dm_c - 2 Dv_c mean0_c (2 - wsum).
      // Two exchange rounds (d var0 rows, then d mean0 rows); A_c / B_c replace mean0 / var0 in the
      // per-sample statistics buffer (each channel is owned by one thread of the sample).
      float d_row[NFB_ROW_CH];
      {
        named_bar_sync(bar_id, GROUP);
#pragma unroll
        for (int c = 0; c < NFB_ROW_CH; ++c) ex[tg * EXS + c] = dot_row<64>(d_h1, sw + W_BASE0 + (35 + c) * 64);
        named_bar_sync(bar_id, GROUP);
        if (active) {
          for (int c = v; c < NFB_ROW_CH; c += V) {
            float dv = 0.f;
            for (int u = 0; u < V; ++u) dv += ex[(base + u) * EXS + c];
            const float m0 = mvs[c];
            mvs[c] = -2.f * dv * m0 * (2.f - wsum0);
            mvs[36 + c] = 2.f * dv;
          }
        }
        named_bar_sync(bar_id, GROUP);
#pragma unroll
        for (int c = 0; c < NFB_ROW_CH; ++c) ex[tg * EXS + c] = dot_row<64>(d_h1, sw + W_BASE0 + c * 64);
        named_bar_sync(bar_id, GROUP);
        if (active) {
          for (int c = v; c < NFB_ROW_CH; c += V) {
            float dm = 0.f;
            for (int u = 0; u < V; ++u) dm += ex[(base + u) * EXS + c];
            mvs[c] += dm;
          }
        }
        named_bar_sync(bar_id, GROUP);
#pragma unroll
        for (int c = 0; c < NFB_ROW_CH; ++c)
          d_row[c] = dot_row<64>(d_h1, sw + W_BASE0 + (70 + c) * 64) + w * (mvs[c] + x[c] * mvs[36 + c]);
        if (WG) {
          // d w_v of the first pooling: sum_c x0_vc A_c + B_c/2 (x0_vc^2 + mean0_c^2)   (A_c, B_c as above)
#pragma unroll
          for (int c = 0; c < NFB_ROW_CH; ++c)
            d_w += x[c] * mvs[c] + 0.5f * mvs[36 + c] * (x[c] * x[c] + m0[c] * m0[c]);
          // ray_dir_fc (:231): d direction_feat = d x0 (d_row before the direct rgb_in term below)
          float a1[16];
          load_bias<16>(a1, sw + B_DIR0);
          dense_acc<4, 16>(sw + W_DIR0, rd, a1);
          elu_inplace<16>(a1);
          float ddf[36];
          load_bias<36>(ddf, sw + B_DIR2);
          dense_acc<16, 36>(sw + W_DIR2, a1, ddf);
#pragma unroll
          for (int c = 0; c < NFB_ROW_CH; ++c) ddf[c] = d_row[c] * elu_grad_from_out(elu_f(ddf[c]));
          ddf[35] = 0.f;
          nfbwg::wgrad_tc<16, 36>(wg, WC_DIR2, a1, ddf, gate);
          wgrad_rowsum<36>(sg + SG_DIR2_B, ddf, gate);
          float da1[16];
          dense_T<16, 36>(sw + W_DIR2, ddf, da1);
#pragma unroll
          for (int k = 0; k < 16; ++k) da1[k] *= elu_grad_from_out(a1[k]);
          nfbwg::wgrad_tc<4, 16>(wg, WC_DIR0, rd, da1, gate);
          wgrad_rowsum<16>(sg + SG_DIR0_B, da1, gate);
        }
        // rgb_in enters the blend directly (:233,272)
        d_row[0] = fmaf(blend, d_r0, d_row[0]);
        d_row[1] = fmaf(blend, d_r1, d_row[1]);
        d_row[2] = fmaf(blend, d_r2, d_row[2]);
        named_bar_sync(bar_id, GROUP);
        if (WG && a.anti_alias) {
          // s (:236-238): w_v = n_v / (sum_u n_u + 1e-8), n_v = (e_v - min_u e_u) mask_v, e_v = exp(|s| (dot_v - 1))
          const float ge = e_aa * (rd[3] - 1.f);   // d e_v / d |s|
          ex[tg * EXS + 0] = d_w;
          ex[tg * EXS + 1] = w;
          ex[tg * EXS + 2] = e_aa;
          ex[tg * EXS + 3] = ge;
          named_bar_sync(bar_id, GROUP);
          float dww = 0.f, emin = 3.4e38f, gmin = 0.f;
          for (int u = 0; u < V; ++u) {
            dww = fmaf(ex[(base + u) * EXS + 0], ex[(base + u) * EXS + 1], dww);
            const float eu = ex[(base + u) * EXS + 2];
            if (eu < emin) { emin = eu; gmin = ex[(base + u) * EXS + 3]; }
          }
          ds_acc += gate * ((d_w - dww) / den_aa) * mk * (ge - gmin);
          named_bar_sync(bar_id, GROUP);
        }
      }

      // (8) hand the row cotangent on
      if (FUSED) {
        if (active) {
          ViewGeom g;
          g.gx = gx; g.gy = gy;
          scatter_row(g, v, a.H, a.W, a.fh, a.fw, d_row, a.d_feat, a.d_imgs);
        }
      } else {
#pragma unroll
        for (int c = 0; c < NFB_ROW_CH; ++c) ex[tg * EXS + c] = d_row[c];
        named_bar_sync(bar_id, GROUP);
        const size_t row0 = (size_t)tile * TS * V;
        const size_t total = (size_t)a.N * V;
        const int rows_here = (int)((total - row0 < (size_t)(TS * V)) ? (total - row0) : (size_t)(TS * V));
        float* dst = a.d_rgb_feat + row0 * NFB_ROW_CH;
        for (int i = tg; i < rows_here * NFB_ROW_CH; i += GROUP) {
          const int rr = i / NFB_ROW_CH, cc = i - rr * NFB_ROW_CH;
          dst[i] = ex[rr * EXS + cc];
        }
        named_bar_sync(bar_id, GROUP);
      }
    }
  }
  if (WG) {
    if (a.anti_alias) {
      const float t = warp_sum(ds_acc);
      if ((threadIdx.x & 31) == 0) atomicAdd(sg + SG_S, t);
    }
    nfbwg::wg_wait(wg);                                        // this group's last MMA batch has landed in TMEM
    nfbtc::fence_before_sync();
    __syncthreads();
    nfbtc::fence_after_sync();
    float* dp = a.d_params;
    if (threadIdx.x < 128) {
      const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
      nfbwg::wg_flush(wg.tmem, w, l, WC_DIR0, 4, 16, dp + P_DIR0_W);
      nfbwg::wg_flush(wg.tmem, w, l, WC_DIR2, 16, 35, dp + P_DIR2_W);
      nfbwg::wg_flush(wg.tmem, w, l, WC_BASE0, 105, 64, dp + P_BASE0_W);
      nfbwg::wg_flush(wg.tmem, w, l, WC_BASE2, 64, 32, dp + P_BASE2_W);
      nfbwg::wg_flush(wg.tmem, w, l, WC_VIS0, 32, 32, dp + P_VIS0_W);
      nfbwg::wg_flush(wg.tmem, w, l, WC_VIS2, 32, 33, dp + P_VIS2_W);
      nfbwg::wg_flush(wg.tmem, w, l, WC_VISB0, 32, 32, dp + P_VISB0_W);
      nfbwg::wg_flush(wg.tmem, w, l, WC_RGB0, 37, 16, dp + P_RGB0_W);
      nfbwg::wg_flush(wg.tmem, w, l, WC_RGB2, 16, 8, dp + P_RGB2_W);
    }
    flush_vec(dp + P_DIR0_B, sg + SG_DIR0_B, 16, threadIdx.x, blockDim.x);
    flush_vec(dp + P_DIR2_B, sg + SG_DIR2_B, 35, threadIdx.x, blockDim.x);
    flush_vec(dp + P_BASE0_B, sg + SG_BASE0_B, 64, threadIdx.x, blockDim.x);
    flush_vec(dp + P_BASE2_B, sg + SG_BASE2_B, 32, threadIdx.x, blockDim.x);
    flush_vec(dp + P_VIS0_B, sg + SG_VIS0_B, 32, threadIdx.x, blockDim.x);
    flush_vec(dp + P_VIS2_B, sg + SG_VIS2_B, 33, threadIdx.x, blockDim.x);
    flush_vec(dp + P_VISB0_B, sg + SG_VISB0_B, 32, threadIdx.x, blockDim.x);
    flush_vec(dp + P_RGB0_B, sg + SG_RGB0_B, 16, threadIdx.x, blockDim.x);
    flush_vec(dp + P_RGB2_B, sg + SG_RGB2_B, 8, threadIdx.x, blockDim.x);
    flush_vec(dp + P_VISB2_W, sg + SG_VISB2, 32, threadIdx.x, blockDim.x);
    flush_vec(dp + P_VISB2_B, sg + SG_VISB2_B, 1, threadIdx.x, blockDim.x);
    flush_vec(dp + P_RGB4_W, sg + SG_RGB4, 8, threadIdx.x, blockDim.x);
    flush_vec(dp + P_RGB4_B, sg + SG_RGB4_B, 1, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0 && a.anti_alias) {
      const float sv = __ldg(a.params + P_S);
      atomicAdd(dp + P_S, (sv > 0.f ? 1.f : (sv < 0.f ? -1.f : 0.f)) * sg[SG_S]);
    }
    nfbtc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) nfbtc::tmem_dealloc(wg.tmem, 512);
  }
}

constexpr size_t view_smem_bytes(int ng, bool wg = false) {
  return (size_t)(W_TOTAL + 16 * NFB_MAX_VIEWS + 4 + ng * GROUP * EXS + ng * TS_MAX * MVS) * sizeof(float) +
         (wg ? (size_t)SG_TOTAL * sizeof(float) + (size_t)ng * (nfbwg::A_TILE_BYTES + nfbwg::B_TILE_BYTES) + ng * 8 + 16 : 0);
}

template <bool FUSED, bool BWD, int NG, bool WG = false>
int launch_view(const ViewArgs& a, cudaStream_t st, const char* name) {
  const size_t smem = view_smem_bytes(NG, WG);
  cudaError_t e = cudaFuncSetAttribute(k_view_stage<FUSED, BWD, NG, WG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
  const int TS = (GROUP / a.V < TS_MAX) ? GROUP / a.V : TS_MAX;
  const int ntiles = (a.N + TS - 1) / TS;
  int grid = (ntiles + NG - 1) / NG;
  const int cap = nfb_num_sms();
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  k_view_stage<FUSED, BWD, NG, WG><<<grid, GROUP * NG, smem, st>>>(a);
  NFB_CHECK_LAUNCH(name);
  return NFB_OK;
}

}  // namespace nfbview

// one explicit instantiation per translation unit (compile time); defined in nfb_view_inst.cu
int nfb_launch_view_tensor_fwd(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_fused_fwd(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_tensor_bwd(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_fused_bwd(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_tensor_wgrad(const nfbview::ViewArgs& a, cudaStream_t st);   // data + parameter gradients
int nfb_launch_view_fused_wgrad(const nfbview::ViewArgs& a, cudaStream_t st);
