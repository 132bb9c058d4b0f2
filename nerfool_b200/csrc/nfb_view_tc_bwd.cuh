// IBRNet view stage on the 5th-generation tensor cores (tcgen05 + TMEM): data-gradient.
//
// One kernel = forward recompute + backward of everything that lives on (sample, view) rows
// (mlp_network.py:231-258, 267-272), ending in the scatter of the row cotangent into d_feat / d_imgs
// (grid_sampler_2d backward, projection.py:119,123).  Same row mapping as the forward kernel in nfb_view_tc.cuh
// (one thread per row, 128-row groups, cross-view reductions through a per-group exchange buffer).
//
// Backward dense layers dX = dY W reuse the FORWARD weight tiles: the tile of W [N_out][K_in] stored K-major is
// read as the MN-major B operand of a 128 x K_in x N_out MMA (instruction-descriptor bit 16), so no transposed
// copies are kept in shared memory.
//
// The activations the backward needs are not kept in registers (the fp32 form spilt 1.4 KB per thread): ELU
// derivatives are stashed as 16-bit codes (nfb_tc.cuh: elu_stash_*), long-lived ones in spare TMEM columns.
// A group owns 256 TMEM columns (2 groups per CTA, one CTA per SM):
//   [0,64)    D accumulator of every MMA
//   [64,112)  x2 (fp32, 32 cols) + ELU'(x1) codes (16 cols) until base_fc.0's backward needs D = 112 columns
//   [112,176) A operand: hi [112,144), lo [144,176)   (1 pass: [112,168) for the 112-wide base_fc.0 input)
//   [176,212) x0 (fp32, 35 used)
//   [212,244) ELU'(h1) codes (32 cols)
#pragma once
#include "nfb_view_tc.cuh"

namespace nfbvtcb {
using namespace nfbtc;
using nfbview::ViewArgs;
using nfbvtc::B_SET_BYTES;
using nfbvtc::EXS;
using nfbvtc::GROUP;
using nfbvtc::layer_k;
using nfbvtc::layer_n;
using nfbvtc::layer_off;
using nfbvtc::MVP;
using nfbvtc::MVS;
using nfbvtc::TS_MAX;
using namespace nfbvtc;   // layer ids, F_* side-table offsets, load_tile, load_side_tables, pack_pair

constexpr int BNG = 2;       // row groups per CTA
constexpr int BGC = 256;     // TMEM columns per group
constexpr int BC_D = 0, BC_X2 = 64, BC_X1Q = 96, BC_A = 112, BC_ALO = 144, BC_X0 = 176, BC_H1Q = 212;
constexpr int DPS = 101;     // staged cotangent row stride (odd): d_mean[32] d_var[32] d_wmean d_rgb[3] | mean[32]

template <int NPASS>
__host__ __device__ constexpr size_t smem_bytes_bwd() {
  return (size_t)B_SET_BYTES * (NPASS == 3 ? 2 : 1) +
         sizeof(float) * (F_TOTAL + 16 * NFB_MAX_VIEWS + 4 + BNG * GROUP * EXS + BNG * TS_MAX * (MVS + MVP) + BNG * TS_MAX * DPS) +
         BNG * 8 + 16;
}

template <int NPASS>
__device__ __forceinline__ void a_put16(uint32_t tl, int kc, const float (&v)[16]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) pack_pair<NPASS>(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  tmem_st8(tl + BC_A + 8 * kc, hi);
  if (NPASS == 3) tmem_st8(tl + BC_ALO + 8 * kc, lo);
}
template <int NPASS>
__device__ __forceinline__ void a_put_words(uint32_t tl, int kc, const uint32_t (&hi)[8], const uint32_t (&lo)[8]) {
  tmem_st8(tl + BC_A + 8 * kc, hi);
  if (NPASS == 3) tmem_st8(tl + BC_ALO + 8 * kc, lo);
}

// y = ELU(D[col..col+16) + bias)
__device__ __forceinline__ void d_elu16(uint32_t tl, int col, const float* __restrict__ bias, float (&y)[16]) {
  tmem_ld16(tl + BC_D + col, y);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias + j);
    y[j + 0] = elu_fast(y[j + 0] + b.x);
    y[j + 1] = elu_fast(y[j + 1] + b.y);
    y[j + 2] = elu_fast(y[j + 2] + b.z);
    y[j + 3] = elu_fast(y[j + 3] + b.w);
  }
}
__device__ __forceinline__ void d_raw16(uint32_t tl, int col, float (&y)[16]) {
  tmem_ld16(tl + BC_D + col, y);
  tmem_ld_wait();
}
__device__ __forceinline__ void stash_codes16(const float (&y)[16], uint32_t (&q)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) q[j] = elu_stash_pack(y[2 * j], y[2 * j + 1]);
}

// forward MMAs of k-steps [KS0, KS1) of LAYER (A chunk index = ks - KS0)
template <int NPASS, int LAYER, int KS0, int KS1>
__device__ __forceinline__ void issue_fwd(uint32_t tb, uint32_t sB_addr, bool acc0) {
  constexpr int N = layer_n(LAYER);
  constexpr uint32_t idesc = idesc_bf16(128, N);
  const uint32_t bhi = sB_addr + layer_off(LAYER), blo = bhi + B_SET_BYTES;
#pragma unroll
  for (int ks = KS0; ks < KS1; ++ks) {
    const uint64_t dh = smem_desc(bhi + ks * 2 * N * 16, N * 16, 128);
    const uint32_t ah = tb + BC_A + 8 * (ks - KS0);
    mma_ts(tb + BC_D, ah, dh, idesc, acc0 || ks > KS0);
    if (NPASS == 3) {
      const uint64_t dl = smem_desc(blo + ks * 2 * N * 16, N * 16, 128);
      mma_ts(tb + BC_D, tb + BC_ALO + 8 * (ks - KS0), dh, idesc, true);
      mma_ts(tb + BC_D, ah, dl, idesc, true);
    }
  }
}
// backward MMAs dX[128][K_in] = dY[128][N_out] W[N_out][K_in]: MMA N = layer_k, MMA K = layer_n, B read MN-major
// from the forward tile (core matrices adjacent in the MMA-K direction are 128 B apart, in the MMA-N direction
// layer_n * 16 B apart; one K = 16 step = two core matrices = 256 B)
template <int NPASS, int LAYER>
__device__ __forceinline__ void issue_bwd(uint32_t tb, uint32_t sB_addr) {
  constexpr int NO = layer_n(LAYER), KI = layer_k(LAYER);
  constexpr uint32_t idesc = idesc_bf16(128, KI) | (1u << 16);
  const uint32_t bhi = sB_addr + layer_off(LAYER), blo = bhi + B_SET_BYTES;
#pragma unroll
  for (int ks = 0; ks < NO / 16; ++ks) {
    const uint64_t dh = smem_desc(bhi + ks * 256, 128, NO * 16);
    const uint32_t ah = tb + BC_A + 8 * ks;
    mma_ts(tb + BC_D, ah, dh, idesc, ks > 0);
    if (NPASS == 3) {
      const uint64_t dl = smem_desc(blo + ks * 256, 128, NO * 16);
      mma_ts(tb + BC_D, tb + BC_ALO + 8 * ks, dh, idesc, true);
      mma_ts(tb + BC_D, ah, dl, idesc, true);
    }
  }
}

#define NFB_TCB_SYNC_ISSUE(STMT)                                              \
  do {                                                                        \
    tmem_st_wait();                                                           \
    fence_before_sync();                                                      \
    named_bar_sync(bar_id, GROUP);                                            \
    if (tg == 0) {                                                            \
      fence_after_sync();                                                     \
      STMT;                                                                   \
      mma_commit(mbar);                                                       \
    }                                                                         \
  } while (0)
#define NFB_TCB_FWD(LAYER, KS0, KS1, ACC0) NFB_TCB_SYNC_ISSUE((issue_fwd<NPASS, LAYER, KS0, KS1>(tb, sB_addr, ACC0)))
#define NFB_TCB_BWD(LAYER) NFB_TCB_SYNC_ISSUE((issue_bwd<NPASS, LAYER>(tb, sB_addr)))
#define NFB_TCB_WAIT()         \
  do {                         \
    mbar_wait(mbar, phase);    \
    phase ^= 1u;               \
    fence_after_sync();        \
  } while (0)

template <int NPASS, bool FUSED>
__global__ void __launch_bounds__(GROUP * BNG, 1) k_view_tc_bwd(ViewArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sB = smem_raw;
  float* sf = reinterpret_cast<float*>(smem_raw + (size_t)B_SET_BYTES * (NPASS == 3 ? 2 : 1));
  float* s_cam = sf + F_TOTAL;
  float* ex_all = s_cam + (16 * NFB_MAX_VIEWS + 4);
  float* mv_all = ex_all + BNG * GROUP * EXS;
  uint32_t* mvp_all = reinterpret_cast<uint32_t*>(mv_all + BNG * TS_MAX * MVS);
  float* dp_all = reinterpret_cast<float*>(mvp_all + BNG * TS_MAX * MVP);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(dp_all + BNG * TS_MAX * DPS + ((BNG * TS_MAX * DPS) & 1));
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + BNG);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int grp = tid / GROUP, tg = tid % GROUP;
  float* ex = ex_all + (size_t)grp * GROUP * EXS;
  float* mv = mv_all + (size_t)grp * TS_MAX * MVS;
  uint32_t* mvp = mvp_all + (size_t)grp * TS_MAX * MVP;
  float* dpb = dp_all + (size_t)grp * TS_MAX * DPS;
  const int bar_id = 1 + grp;
  uint64_t* mbar = s_bar + grp;

  if (warp == 0) tmem_alloc(s_tmem, BNG * BGC);
  if (tid == 0) {
    for (int g = 0; g < BNG; ++g) mbar_init(s_bar + g, 1);
    mbar_init_fence();
  }
  {
    const float* p = a.params;
    load_tile<NPASS>(sB, L_DIR2, p + P_DIR2_W, 35, 16, tid, blockDim.x);
    load_tile<NPASS>(sB, L_BASE0, p + P_BASE0_W, 64, 105, tid, blockDim.x);
    load_tile<NPASS>(sB, L_BASE2, p + P_BASE2_W, 32, 64, tid, blockDim.x);
    load_tile<NPASS>(sB, L_VIS0, p + P_VIS0_W, 32, 32, tid, blockDim.x);
    load_tile<NPASS>(sB, L_VIS2, p + P_VIS2_W, 33, 32, tid, blockDim.x);
    load_tile<NPASS>(sB, L_VISB0, p + P_VISB0_W, 32, 32, tid, blockDim.x);
    load_tile<NPASS>(sB, L_RGB0, p + P_RGB0_W, 16, 37, tid, blockDim.x);
    load_side_tables(sf, p, tid, blockDim.x);
    if (FUSED)
      for (int i = tid; i < 16 * a.V + 3; i += blockDim.x) s_cam[i] = __ldg(a.cam + i);
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const uint32_t tb = *s_tmem + (uint32_t)(grp * BGC);
  const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t sB_addr = smem_u32(sB);
  uint32_t phase = 0;

  const int V = a.V;
  const int TS = (GROUP / V < TS_MAX) ? GROUP / V : TS_MAX;
  const int sl = tg / V, v = tg - sl * V;
  const int ntiles = (a.N + TS - 1) / TS;
  const float Wm1 = (float)a.W - 1.f, Hm1 = (float)a.H - 1.f;
  const float s_abs = sf[F_S];

  for (int tile = blockIdx.x * BNG + grp; tile < ntiles; tile += gridDim.x * BNG) {
    const int p = tile * TS + sl;
    const bool active = (sl < TS) && (p < a.N);
    const int base = active ? sl * V : 0;
    float* mvs = mv + (active ? sl : 0) * MVS;
    uint32_t* mvps = mvp + (active ? sl : 0) * MVP;
    const float* dp = dpb + (active ? sl : 0) * DPS;

    // ---------------- stage the cotangents / forward means of the tile's samples ----------------
    {
      const int p0 = tile * TS;
      const int ns = (a.N - p0 < TS) ? (a.N - p0) : TS;
      // 17 float4 of d_ps (68 floats) + 8 float4 of ps (mean) per sample
      for (int i = tg; i < ns * 25; i += GROUP) {
        const int s = i / 25, j = i - s * 25;
        if (j < 17) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(a.d_ps + (size_t)(p0 + s) * NFB_PS_STRIDE) + j);
          float* d = dpb + s * DPS + 4 * j;
          d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
        } else {
          const float4 q = __ldg(reinterpret_cast<const float4*>(a.ps + (size_t)(p0 + s) * NFB_PS_STRIDE) + (j - 17));
          float* d = dpb + s * DPS + 68 + 4 * (j - 17);
          d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
        }
      }
    }

    // ---------------- projection, ray_diff, bilinear gather ----------------
    float x[NFB_ROW_CH];
    float rd[4];
    float mk = 0.f, gx = 0.f, gy = 0.f;
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
        const size_t row = (size_t)p * V + v;
        const float4 q = __ldg(reinterpret_cast<const float4*>(a.ray_diff) + row);
        rd[0] = q.x; rd[1] = q.y; rd[2] = q.z; rd[3] = q.w;
        mk = __ldg(a.mask + row);
      } else {
        rd[0] = rd[1] = rd[2] = rd[3] = 0.f;
      }
      named_bar_sync(bar_id, GROUP);
    }
    const float rgb_in0 = x[0], rgb_in1 = x[1], rgb_in2 = x[2];

    // ---------------- ray_dir_fc ----------------
    {
      float a1[16];
      load_bias<16>(a1, sf + F_DIR0_B);
      dense_acc<4, 16>(sf + F_DIR0_W, rd, a1);
#pragma unroll
      for (int j = 0; j < 16; ++j) a1[j] = elu_fast(a1[j]);
      a_put16<NPASS>(tl, 0, a1);
    }
    NFB_TCB_FWD(L_DIR2, 0, 1, false);

    // ---------------- pooling weights (overlaps the MMA) ----------------
    float w, wsum0;
    {
      const float e = a.anti_alias ? (float)exp((double)__fmul_rn(s_abs, __fsub_rn(rd[3], 1.f))) : 1.f;
      ex[tg * EXS + 35] = e;
      ex[tg * EXS + 36] = mk;
      named_bar_sync(bar_id, GROUP);
      float mn = 3.4e38f;
      for (int u = 0; u < V; ++u) mn = fminf(mn, ex[(base + u) * EXS + 35]);
      if (!a.anti_alias) mn = 0.f;
      float sum = 0.f;
      for (int u = 0; u < V; ++u) sum += (ex[(base + u) * EXS + 35] - mn) * ex[(base + u) * EXS + 36];
      const float inv = 1.f / (sum + 1e-8f);
      w = (e - mn) * mk * inv;
      wsum0 = 0.f;
      for (int u = 0; u < V; ++u) wsum0 += (ex[(base + u) * EXS + 35] - mn) * ex[(base + u) * EXS + 36] * inv;
      named_bar_sync(bar_id, GROUP);
    }

    // ---------------- x0 = rgb_feat + direction_feat ; stash x0 ----------------
    NFB_TCB_WAIT();
#pragma unroll
    for (int c0 = 0; c0 < 48; c0 += 16) {
      float df[16];
      d_elu16(tl, c0, sf + F_B_DIR2 + c0, df);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c0 + j < NFB_ROW_CH) x[c0 + j] += df[j];
    }
    {
      uint32_t t16[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) t16[j] = __float_as_uint(x[j]);
      tmem_st16(tl + BC_X0, t16);
#pragma unroll
      for (int j = 0; j < 16; ++j) t16[j] = __float_as_uint(x[16 + j]);
      tmem_st16(tl + BC_X0 + 16, t16);
      uint32_t t4[4] = {__float_as_uint(x[32]), __float_as_uint(x[33]), __float_as_uint(x[34]), 0u};
      tmem_st4(tl + BC_X0 + 32, t4);
    }

    // ---------------- first pooling ----------------
#pragma unroll
    for (int c = 0; c < NFB_ROW_CH; ++c) ex[tg * EXS + c] = x[c];
    ex[tg * EXS + 35] = w;
    named_bar_sync(bar_id, GROUP);
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
        __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(mvps);
        __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(mvps + 36);
        const __nv_bfloat16 mh = __float2bfloat16_rn(m), vh = __float2bfloat16_rn(vr);
        ph[c] = mh;
        ph[35 + c] = vh;
        if (NPASS == 3) {
          pl[c] = __float2bfloat16_rn(m - __bfloat162float(mh));
          pl[35 + c] = __float2bfloat16_rn(vr - __bfloat162float(vh));
        }
      }
    }
    named_bar_sync(bar_id, GROUP);

    // ---------------- base_fc.0 ----------------
    {
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        uint32_t hi[8], lo[8];
        const uint4 h0 = *reinterpret_cast<const uint4*>(mvps + 8 * kc), h1 = *reinterpret_cast<const uint4*>(mvps + 8 * kc + 4);
        hi[0] = h0.x; hi[1] = h0.y; hi[2] = h0.z; hi[3] = h0.w; hi[4] = h1.x; hi[5] = h1.y; hi[6] = h1.z; hi[7] = h1.w;
        if (NPASS == 3) {
          const uint4 l0 = *reinterpret_cast<const uint4*>(mvps + 36 + 8 * kc), l1 = *reinterpret_cast<const uint4*>(mvps + 36 + 8 * kc + 4);
          lo[0] = l0.x; lo[1] = l0.y; lo[2] = l0.z; lo[3] = l0.w; lo[4] = l1.x; lo[5] = l1.y; lo[6] = l1.z; lo[7] = l1.w;
        }
        a_put_words<NPASS>(tl, kc, hi, lo);
      }
      if (NPASS == 3) {
        NFB_TCB_FWD(L_BASE0, 0, 4, false);
        NFB_TCB_WAIT();
      }
      constexpr int KC0 = (NPASS == 3) ? 4 : 0;
      {
        uint32_t hi[8], lo[8];
        hi[0] = mvps[32]; hi[1] = mvps[33]; hi[2] = mvps[34];
        if (NPASS == 3) { lo[0] = mvps[36 + 32]; lo[1] = mvps[36 + 33]; lo[2] = mvps[36 + 34]; }
#pragma unroll
        for (int j = 0; j < 5; ++j) pack_pair<NPASS>(x[2 * j], x[2 * j + 1], hi[3 + j], lo[3 + j]);
        a_put_words<NPASS>(tl, 4 - KC0, hi, lo);
#pragma unroll
        for (int j = 0; j < 8; ++j) pack_pair<NPASS>(x[10 + 2 * j], x[11 + 2 * j], hi[j], lo[j]);
        a_put_words<NPASS>(tl, 5 - KC0, hi, lo);
#pragma unroll
        for (int j = 0; j < 4; ++j) pack_pair<NPASS>(x[26 + 2 * j], x[27 + 2 * j], hi[j], lo[j]);
        pack_pair<NPASS>(x[34], 0.f, hi[4], lo[4]);
        hi[5] = hi[6] = hi[7] = 0u;
        lo[5] = lo[6] = lo[7] = 0u;
        a_put_words<NPASS>(tl, 6 - KC0, hi, lo);
      }
      if (NPASS == 3) NFB_TCB_FWD(L_BASE0, 4, 7, true);
      else NFB_TCB_FWD(L_BASE0, 0, 7, false);
      NFB_TCB_WAIT();
    }

    // ---------------- base_fc.2 ----------------
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      float h[16];
      d_elu16(tl, 16 * kc, sf + F_B_BASE0 + 16 * kc, h);
      uint32_t q[8];
      stash_codes16(h, q);
      tmem_st8(tl + BC_H1Q + 8 * kc, q);
      a_put16<NPASS>(tl, kc, h);
    }
    NFB_TCB_FWD(L_BASE2, 0, 4, false);
    NFB_TCB_WAIT();

    // ---------------- vis_fc ----------------
    float x1[32];
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      float h[16];
      d_elu16(tl, 16 * kc, sf + F_B_BASE2 + 16 * kc, h);
      uint32_t q[8];
      stash_codes16(h, q);
      tmem_st8(tl + BC_X1Q + 8 * kc, q);
      float t[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        x1[16 * kc + j] = h[j];
        t[j] = h[j] * w;
      }
      a_put16<NPASS>(tl, kc, t);
    }
    NFB_TCB_FWD(L_VIS0, 0, 2, false);
    NFB_TCB_WAIT();
    uint32_t hvq[16];
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      float h[16];
      d_elu16(tl, 16 * kc, sf + F_B_VIS0 + 16 * kc, h);
#pragma unroll
      for (int j = 0; j < 8; ++j) hvq[8 * kc + j] = elu_stash_pack(h[2 * j], h[2 * j + 1]);
      a_put16<NPASS>(tl, kc, h);
    }
    NFB_TCB_FWD(L_VIS2, 0, 2, false);
    NFB_TCB_WAIT();

    uint32_t xvq[17];
    float sg1, vis1;
    {
      float h[16];
      d_elu16(tl, 32, sf + F_B_VIS2 + 32, h);
      xvq[16] = elu_stash_pack(h[0], 0.f);
      sg1 = sigmoid_f(h[0]);
      vis1 = sg1 * mk;
    }
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      float h[16];
      d_elu16(tl, 16 * kc, sf + F_B_VIS2 + 16 * kc, h);
#pragma unroll
      for (int j = 0; j < 8; ++j) xvq[8 * kc + j] = elu_stash_pack(h[2 * j], h[2 * j + 1]);
      float t[16];
      uint32_t xs[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        x1[16 * kc + j] += h[j];            // x1 now holds x2
        xs[j] = __float_as_uint(x1[16 * kc + j]);
        t[j] = x1[16 * kc + j] * vis1;
      }
      tmem_st16(tl + BC_X2 + 16 * kc, xs);
      a_put16<NPASS>(tl, kc, t);
    }
    NFB_TCB_FWD(L_VISB0, 0, 2, false);
    NFB_TCB_WAIT();
    uint32_t hv2q[16];
    float sg2, vis2;
    {
      float z = sf[F_B_VISB2];
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        float h[16];
        d_elu16(tl, 16 * kc, sf + F_B_VISB0 + 16 * kc, h);
#pragma unroll
        for (int j = 0; j < 8; ++j) hv2q[8 * kc + j] = elu_stash_pack(h[2 * j], h[2 * j + 1]);
#pragma unroll
        for (int j = 0; j < 16; ++j) z = fmaf(h[j], sf[F_W_VISB2 + 16 * kc + j], z);
      }
      sg2 = sigmoid_f(z);
      vis2 = sg2 * mk;
    }

    // ---------------- rgb_fc ----------------
    {
      float t[16];
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
#pragma unroll
        for (int j = 0; j < 16; ++j) t[j] = x1[16 * kc + j];
        a_put16<NPASS>(tl, kc, t);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) t[j] = 0.f;
      t[0] = vis2; t[1] = rd[0]; t[2] = rd[1]; t[3] = rd[2]; t[4] = rd[3];
      a_put16<NPASS>(tl, 2, t);
    }
    NFB_TCB_FWD(L_RGB0, 0, 3, false);
    NFB_TCB_WAIT();
    float g1d[16], g2d[8];     // ELU' of the two hidden layers of rgb_fc
    float logit;
    {
      float g1[16];
      d_elu16(tl, 0, sf + F_B_RGB0, g1);
      float g2[8];
      load_bias<8>(g2, sf + F_B_RGB2);
      dense_acc<16, 8>(sf + F_W_RGB2, g1, g2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g2[j] = elu_fast(g2[j]);
        g2d[j] = elu_grad_from_out(g2[j]);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) g1d[j] = elu_grad_from_out(g1[j]);
      logit = dot_row<8>(g2, sf + F_W_RGB4) + sf[F_B_RGB4];
      if (mk == 0.f) logit = -1e9f;
    }

    // ---------------- exchange vis2 / logit / rgb_in ; blending softmax ----------------
    ex[tg * EXS + 32] = vis2;
    ex[tg * EXS + 33] = logit;
    ex[tg * EXS + 34] = rgb_in0;
    ex[tg * EXS + 35] = rgb_in1;
    ex[tg * EXS + 36] = rgb_in2;
    named_bar_sync(bar_id, GROUP);
    float Dsum = 1e-8f;
    for (int u = 0; u < V; ++u) Dsum += ex[(base + u) * EXS + 32];
    const float invD = 1.f / Dsum;
    float mx = -3.4e38f;
    for (int u = 0; u < V; ++u) mx = fmaxf(mx, ex[(base + u) * EXS + 33]);
    float se = 0.f;
    for (int u = 0; u < V; ++u) se += __expf(ex[(base + u) * EXS + 33] - mx);
    const float inv_se = 1.f / se;

    // =================================== backward ===================================
    const float w2 = vis2 * invD;
    float w2sum = 0.f;
    for (int u = 0; u < V; ++u) w2sum += ex[(base + u) * EXS + 32] * invD;
    const float d_r0 = dp[65], d_r1 = dp[66], d_r2 = dp[67];
    const float d_wmean = dp[64];

    // (1) blending softmax
    const float blend = __expf(logit - mx) * inv_se;
    float d_logit;
    {
      float bt = 0.f;
      for (int u = 0; u < V; ++u) {
        const float b = __expf(ex[(base + u) * EXS + 33] - mx) * inv_se;
        const float tu = ex[(base + u) * EXS + 34] * d_r0 + ex[(base + u) * EXS + 35] * d_r1 + ex[(base + u) * EXS + 36] * d_r2;
        bt = fmaf(b, tu, bt);
      }
      const float tv = rgb_in0 * d_r0 + rgb_in1 * d_r1 + rgb_in2 * d_r2;
      d_logit = (mk != 0.f) ? blend * (tv - bt) : 0.f;
    }

    // (2) rgb_fc backward: the two small layers on the CUDA cores, rgb_fc.0 as MMA (N = 48: d[x2 | vis2 | ray_diff])
    {
      float dg2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dg2[j] = d_logit * sf[F_W_RGB4 + j] * g2d[j];
      float dg1[16];
      dense_T<16, 8>(sf + F_W_RGB2, dg2, dg1);
#pragma unroll
      for (int k = 0; k < 16; ++k) dg1[k] *= g1d[k];
      a_put16<NPASS>(tl, 0, dg1);
    }
    NFB_TCB_BWD(L_RGB0);

    // (3) second pooling backward (overlaps the MMA)
    float x2[32];
    float d_x2[32];
    float d_w2 = d_wmean / (float)V;
    {
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        uint32_t xs[16];
        tmem_ld16u(tl + BC_X2 + 16 * kc, xs);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) x2[16 * kc + j] = __uint_as_float(xs[j]);
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float dm = dp[c], dv = dp[32 + c];
        const float mean = dp[68 + c];
        const float diff = x2[c] - mean;
        d_w2 = fmaf(dm, x2[c], d_w2);
        d_w2 = fmaf(dv * diff, diff, d_w2);
        d_x2[c] = w2 * (dm - 2.f * dv * mean * (1.f - w2sum)) + 2.f * w2 * diff * dv;
      }
    }
    named_bar_sync(bar_id, GROUP);          // all reads of slots 33..36 above are done
    ex[tg * EXS + 33] = d_w2 * vis2;
    named_bar_sync(bar_id, GROUP);
    float d_vis2;
    {
      float sdv = 0.f;
      for (int u = 0; u < V; ++u) sdv += ex[(base + u) * EXS + 33];
      d_vis2 = d_w2 * invD - sdv * invD * invD;
    }
    NFB_TCB_WAIT();
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      float t[16];
      d_raw16(tl, 16 * kc, t);
#pragma unroll
      for (int j = 0; j < 16; ++j) d_x2[16 * kc + j] += t[j];
    }
    {
      float t[16];
      d_raw16(tl, 32, t);
      d_vis2 += t[0];
    }

    // (4) vis_fc2 backward
    float d_vis1;
    {
      const float dz = d_vis2 * mk * sg2 * (1.f - sg2);
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        float dh[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dh[2 * j] = dz * sf[F_W_VISB2 + 16 * kc + 2 * j] * elu_stash_lo(hv2q[8 * kc + j]);
          dh[2 * j + 1] = dz * sf[F_W_VISB2 + 16 * kc + 2 * j + 1] * elu_stash_hi(hv2q[8 * kc + j]);
        }
        a_put16<NPASS>(tl, kc, dh);
      }
      NFB_TCB_BWD(L_VISB0);
      NFB_TCB_WAIT();
      d_vis1 = 0.f;
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        float dt[16];
        d_raw16(tl, 16 * kc, dt);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          d_vis1 = fmaf(dt[j], x2[16 * kc + j], d_vis1);
          d_x2[16 * kc + j] = fmaf(dt[j], vis1, d_x2[16 * kc + j]);
        }
      }
    }

    // (5) vis_fc backward: d xv = [d_x2 | d_vis1 path] * ELU'(xv)
    {
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        float dxv[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dxv[2 * j] = d_x2[16 * kc + 2 * j] * elu_stash_lo(xvq[8 * kc + j]);
          dxv[2 * j + 1] = d_x2[16 * kc + 2 * j + 1] * elu_stash_hi(xvq[8 * kc + j]);
        }
        a_put16<NPASS>(tl, kc, dxv);
      }
      float dxv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) dxv[j] = 0.f;
      dxv[0] = d_vis1 * mk * sg1 * (1.f - sg1) * elu_stash_lo(xvq[16]);
      a_put16<NPASS>(tl, 2, dxv);
    }
    NFB_TCB_BWD(L_VIS2);
    NFB_TCB_WAIT();
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      float dh[16];
      d_raw16(tl, 16 * kc, dh);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dh[2 * j] *= elu_stash_lo(hvq[8 * kc + j]);
        dh[2 * j + 1] *= elu_stash_hi(hvq[8 * kc + j]);
      }
      a_put16<NPASS>(tl, kc, dh);
    }
    NFB_TCB_BWD(L_VIS0);
    NFB_TCB_WAIT();

    // (6) base_fc backward
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      float dt[16];
      d_raw16(tl, 16 * kc, dt);
      uint32_t q[8];
      tmem_ld8u(tl + BC_X1Q + 8 * kc, q);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dt[2 * j] = fmaf(dt[2 * j], w, d_x2[16 * kc + 2 * j]) * elu_stash_lo(q[j]);
        dt[2 * j + 1] = fmaf(dt[2 * j + 1], w, d_x2[16 * kc + 2 * j + 1]) * elu_stash_hi(q[j]);
      }
      a_put16<NPASS>(tl, kc, dt);
    }
    NFB_TCB_BWD(L_BASE2);
    NFB_TCB_WAIT();
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      float dh[16];
      d_raw16(tl, 16 * kc, dh);
      uint32_t q[8];
      tmem_ld8u(tl + BC_H1Q + 8 * kc, q);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dh[2 * j] *= elu_stash_lo(q[j]);
        dh[2 * j + 1] *= elu_stash_hi(q[j]);
      }
      a_put16<NPASS>(tl, kc, dh);
    }
    NFB_TCB_BWD(L_BASE0);                     // D = [d mean0 (35) | d var0 (35) | d x0 (35) | pad] in columns [0,112)
    NFB_TCB_WAIT();

    // (7) first pooling backward.  With Dm_c = sum_v d mean0_vc, Dv_c = sum_v d var0_vc:
    //   d x0_vc = dx_vc + w_v A_c + w_v x0_vc B_c,  B_c = 2 Dv_c,  A_c = Dm_c - 2 Dv_c mean0_c (2 - wsum)
    float d_row[NFB_ROW_CH];
    {
      // d var0: columns 35..69
#pragma unroll
      for (int c0 = 32; c0 < 80; c0 += 16) {
        float t[16];
        d_raw16(tl, c0, t);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j >= 35 && c0 + j < 70) ex[tg * EXS + (c0 + j - 35)] = t[j];
      }
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
      // d mean0: columns 0..34
#pragma unroll
      for (int c0 = 0; c0 < 48; c0 += 16) {
        float t[16];
        d_raw16(tl, c0, t);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < 35) ex[tg * EXS + c0 + j] = t[j];
      }
      named_bar_sync(bar_id, GROUP);
      if (active) {
        for (int c = v; c < NFB_ROW_CH; c += V) {
          float dm = 0.f;
          for (int u = 0; u < V; ++u) dm += ex[(base + u) * EXS + c];
          mvs[c] += dm;
        }
      }
      named_bar_sync(bar_id, GROUP);
      // d x0: columns 70..104, + pooled terms
      float x0[36];
      {
        uint32_t t16[16];
        tmem_ld16u(tl + BC_X0, t16);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) x0[j] = __uint_as_float(t16[j]);
        tmem_ld16u(tl + BC_X0 + 16, t16);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) x0[16 + j] = __uint_as_float(t16[j]);
        uint32_t t4[4];
        tmem_ld4u(tl + BC_X0 + 32, t4);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) x0[32 + j] = __uint_as_float(t4[j]);
      }
#pragma unroll
      for (int c0 = 64; c0 < 112; c0 += 16) {
        float t[16];
        d_raw16(tl, c0, t);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = c0 + j - 70;
          if (c >= 0 && c < NFB_ROW_CH) d_row[c] = t[j] + w * (mvs[c] + x0[c] * mvs[36 + c]);
        }
      }
      d_row[0] = fmaf(blend, d_r0, d_row[0]);
      d_row[1] = fmaf(blend, d_r1, d_row[1]);
      d_row[2] = fmaf(blend, d_r2, d_row[2]);
    }

    // (8) hand the row cotangent on
    if (FUSED) {
      if (active) {
        ViewGeom g;
        g.gx = gx; g.gy = gy;
        scatter_row(g, v, a.H, a.W, a.fh, a.fw, d_row, a.d_feat, a.d_imgs);
      }
      named_bar_sync(bar_id, GROUP);          // exchange / statistics buffers are reused by the next tile
    } else {
      named_bar_sync(bar_id, GROUP);
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

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*s_tmem, BNG * BGC);
}

template <int NPASS, bool FUSED>
int launch_view_tc_bwd(const ViewArgs& a, cudaStream_t st) {
  constexpr size_t smem = smem_bytes_bwd<NPASS>();
  cudaError_t e = cudaFuncSetAttribute(k_view_tc_bwd<NPASS, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "k_view_tc_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int TS = (GROUP / a.V < TS_MAX) ? GROUP / a.V : TS_MAX;
  const int ntiles = (a.N + TS - 1) / TS;
  int grid = (ntiles + BNG - 1) / BNG;
  const int cap = nfb_num_sms();
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  k_view_tc_bwd<NPASS, FUSED><<<grid, GROUP * BNG, smem, st>>>(a);
  NFB_CHECK_LAUNCH("k_view_tc_bwd");
  return NFB_OK;
}

}  // namespace nfbvtcb

int nfb_launch_view_tc_bwd_p1_fused(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_tc_bwd_p3_fused(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_tc_bwd_p1_tensor(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_tc_bwd_p3_tensor(const nfbview::ViewArgs& a, cudaStream_t st);
