// IBRNet view stage, forward, fused form -- TWO threads per (sample, view) row.
//
// Same mathematics, tile mapping, TMEM operand scheme, weight tiles and activation-stash layout as k_view_tc_fwd
// (nfb_view_tc.cuh), so the stash backward (nfb_view_tc_bwd2.cuh) reads what this kernel writes.  What changes is the
// occupancy: k_view_tc_fwd is latency bound at 16 warps per SM (4 groups x 128 threads x 128 registers; TMEM = 4 x 128
// columns caps the number of groups).  Here a 128-row group is served by 256 threads: warps q and q + 4 of a group
// address the SAME 32 TMEM lanes (a warp may touch lanes 32 (warp % 4) .. +31), thread (row, half) owns half of the
// row's columns in every epilogue, half of the gathered channels and half of every A-operand chunk.  3 groups x 256
// threads = 24 warps per SM at <= 85 registers, TMEM 3 x 128 columns.
//   half 0 ("A"): RGB + feature channels 0..15 (x[0..19)), output columns [0, N/2) of every layer, the scalar tail (rgb_fc.2/.4)
//   half 1 ("B"): feature channels 16..31 (x[19..35)), output columns [N/2, N)
// Cross-view exchanges: the rows of a sample sit in warp q (A halves) and warp q + 4 (B halves): a 64-thread named barrier per
// warp pair replaces the warp-level fence of the one-thread form (16 hardware barriers: 0 | 3 groups | 12 pairs).
// base_fc.0's K dimension is permuted so that both halves write whole, aligned column blocks:
//   k' 0..69 [mean | var], 70..71 zero, 72..90 x[0..19), 91..95 zero, 96..111 x[19..35).
#pragma once
#include "nfb_view_tc.cuh"

namespace nfbvtc2 {
using namespace nfbtc;
using namespace nfbvtc;
using nfbview::ViewArgs;

constexpr int NG2 = 3;
constexpr int GT = 2 * GROUP;            // threads per group
constexpr int XH = 20;                   // per-thread half row (A: 19 values + pad, B: 16 values)
enum : int { F2_B_DIR2B = F_TOTAL, F2_TOTAL = F_TOTAL + 16 };     // + ray_dir_fc.2 bias [19..35) on a 16-byte boundary (half B)

// source column of base_fc.0.weight [64][105] for tile column k' (-1 = zero)
__host__ __device__ constexpr int base0_src(int k) {
  return k < 70 ? k : k < 72 ? -1 : k < 91 ? 70 + (k - 72) : k < 96 ? -1 : 70 + 19 + (k - 96);
}

template <int NPASS>
__host__ __device__ constexpr size_t smem_bytes2() {
  return (size_t)B_SET_BYTES * (NPASS == 3 ? 2 : 1) +
         sizeof(float) * (F2_TOTAL + 16 * NFB_MAX_VIEWS + 4 + NG2 * GROUP * (EXQ + SIDE) + NG2 * TS_MAX * MVP) + NG2 * 8 + 16;
}

template <int NPASS>
static __device__ void load_tile_base0_perm(uint8_t* sB, const float* __restrict__ w, int tid, int nt) {
  constexpr int N = layer_n(L_BASE0), K = layer_k(L_BASE0);
  uint8_t* hi = sB + layer_off(L_BASE0);
  uint8_t* lo = hi + B_SET_BYTES;
  for (int i = tid; i < N * K; i += nt) {
    const int n = i / K, k = i - n * K;
    const int src = base0_src(k);
    const float v = src >= 0 ? __ldg(w + n * 105 + src) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const uint32_t off = canon_off(n, k, N);
    *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
    if (NPASS == 3) *reinterpret_cast<__nv_bfloat16*>(lo + off) = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// half of the gathered row: A = RGB + feature channels 0..15 -> xh[0..19), B = feature channels 16..31 -> xh[0..16)
template <int HALF>
__device__ __forceinline__ void gather_half(const ViewGeom& g, int v, int H, int W, int fh, int fw, const float* __restrict__ imgs,
                                            const float* __restrict__ feat, float (&xh)[XH]) {
#pragma unroll
  for (int c = 0; c < XH; ++c) xh[c] = 0.f;
  if (HALF == 0) {
    const Taps t = bilinear_taps(g.gx, g.gy, W, H);
    const float* base = imgs + (size_t)v * H * W * 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (t.off[i] >= 0) {
        const float* p = base + (size_t)t.off[i] * 3;
        xh[0] += __ldg(p + 0) * t.wt[i];
        xh[1] += __ldg(p + 1) * t.wt[i];
        xh[2] += __ldg(p + 2) * t.wt[i];
      }
    }
  }
  constexpr int O = HALF == 0 ? 3 : 0;
  const Taps t = bilinear_taps(g.gx, g.gy, fw, fh);
  const float4* base = reinterpret_cast<const float4*>(feat + (size_t)v * fh * fw * NFB_FEAT_CH) + HALF * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (t.off[i] >= 0) {
      const float4* p = base + (size_t)t.off[i] * (NFB_FEAT_CH / 4);
      const float wgt = t.wt[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 q = __ldg(p + j);
        xh[O + 4 * j + 0] += q.x * wgt;
        xh[O + 4 * j + 1] += q.y * wgt;
        xh[O + 4 * j + 2] += q.z * wgt;
        xh[O + 4 * j + 3] += q.w * wgt;
      }
    }
  }
}

// cross-view pooling on float4 quads with an explicit start / stride (pool4 of nfb_view_tc.cuh: q = v, v + V, ...)
template <int NQ, int MODE, typename EMIT>
__device__ __forceinline__ void pool4x(const float* __restrict__ row0, int V, int q0, int qs, int wslot, float scale, EMIT&& emit) {
  for (int q = q0; q < NQ; q += qs) {
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int u = 0; u < V; ++u) {
      const float* r = row0 + u * EXQ;
      const float4 x = *reinterpret_cast<const float4*>(r + 4 * q);
      const float wu = r[wslot] * scale;
      m.x = fmaf(x.x, wu, m.x); m.y = fmaf(x.y, wu, m.y); m.z = fmaf(x.z, wu, m.z); m.w = fmaf(x.w, wu, m.w);
    }
    float4 s2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == POOL_MEAN_VAR) {
#pragma unroll 2
      for (int u = 0; u < V; ++u) {
        const float* r = row0 + u * EXQ;
        const float4 x = *reinterpret_cast<const float4*>(r + 4 * q);
        const float wu = r[wslot] * scale;
        const float dx = x.x - m.x, dy = x.y - m.y, dz = x.z - m.z, dw = x.w - m.w;
        s2.x = fmaf(wu * dx, dx, s2.x); s2.y = fmaf(wu * dy, dy, s2.y); s2.z = fmaf(wu * dz, dz, s2.z); s2.w = fmaf(wu * dw, dw, s2.w);
      }
    }
    emit(q, m, s2);
  }
}

__device__ __forceinline__ void tmem_st16w(uint32_t taddr, const uint32_t (&r)[16]) { tmem_st16(taddr, r); }

#define NFB_TC2_ISSUE(LAYER, KS0, KS1, ACC0, ISSUER)                          \
  do {                                                                        \
    tmem_st_wait();                                                           \
    fence_before_sync();                                                      \
    named_bar_sync(bar_id, GT);                                               \
    if (tg == 32 * (ISSUER)) {                                                \
      fence_after_sync();                                                     \
      issue_mma<NPASS, LAYER, KS0, KS1>(tb, sB_addr, ACC0);                   \
      mma_commit(mbar);                                                       \
    }                                                                         \
  } while (0)
#define NFB_EX2_SYNC()                                  \
  do {                                                  \
    if (warp_local) named_bar_sync(pair_id, 64);        \
    else named_bar_sync(bar_id, GT);                    \
  } while (0)

template <int NPASS, bool SAVE>
__global__ void __launch_bounds__(GT * NG2, 1) k_view_tc_fwd2(ViewArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sB = smem_raw;
  float* sf = reinterpret_cast<float*>(smem_raw + (size_t)B_SET_BYTES * (NPASS == 3 ? 2 : 1));
  float* s_cam = sf + F2_TOTAL;
  float* ex_all = s_cam + (16 * NFB_MAX_VIEWS + 4);
  float* side_all = ex_all + NG2 * GROUP * EXQ;
  uint32_t* mvp_all = reinterpret_cast<uint32_t*>(side_all + NG2 * GROUP * SIDE);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(mvp_all + NG2 * TS_MAX * MVP);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + NG2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int grp = tid / GT, tg = tid % GT;
  const int half = tg >> 7, row = tg & (GROUP - 1);
  float* ex = ex_all + (size_t)grp * GROUP * EXQ;
  float* side = side_all + (size_t)grp * GROUP * SIDE;
  uint32_t* mvp = mvp_all + (size_t)grp * TS_MAX * MVP;
  const int bar_id = 1 + grp;
  const int pair_id = 1 + NG2 + grp * 4 + (warp & 3);
  uint64_t* mbar = s_bar + grp;

  if (warp == 0) tmem_alloc(s_tmem, 512);
  if (tid == 0) {
    for (int g = 0; g < NG2; ++g) mbar_init(s_bar + g, 1);
    mbar_init_fence();
  }
  {
    const float* p = a.params;
    load_tile<NPASS>(sB, L_DIR2, p + P_DIR2_W, 35, 16, tid, blockDim.x);
    load_tile_base0_perm<NPASS>(sB, p + P_BASE0_W, tid, blockDim.x);
    load_tile<NPASS>(sB, L_BASE2, p + P_BASE2_W, 32, 64, tid, blockDim.x);
    load_tile<NPASS>(sB, L_VIS0, p + P_VIS0_W, 32, 32, tid, blockDim.x);
    load_tile<NPASS>(sB, L_VIS2, p + P_VIS2_W, 33, 32, tid, blockDim.x);
    load_tile<NPASS>(sB, L_VISB0, p + P_VISB0_W, 32, 32, tid, blockDim.x);
    load_tile<NPASS>(sB, L_RGB0, p + P_RGB0_W, 16, 37, tid, blockDim.x);
    load_side_tables(sf, p, tid, blockDim.x);
    for (int i = tid; i < 16; i += blockDim.x) sf[F2_B_DIR2B + i] = __ldg(p + P_DIR2_B + 19 + i);
    for (int i = tid; i < 16 * a.V + 3; i += blockDim.x) s_cam[i] = __ldg(a.cam + i);
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const uint32_t tb = *s_tmem + (uint32_t)(grp * GC);
  const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t sB_addr = smem_u32(sB);
  uint32_t phase = 0;

  const int V = a.V;
  const RowMap rm = row_map(V);
  const bool warp_local = rm.packed;
  const int TS = rm.TS;
  int sl, v;
  bool lane_ok;
  if (rm.packed) {
    const int wq = row >> 5, l = row & 31, si = l / V;
    v = l - si * V; sl = wq * rm.spw + si; lane_ok = si < rm.spw;
  } else {
    sl = row / V; v = row - sl * V; lane_ok = sl < TS;
  }
  const int ntiles = (a.N + TS - 1) / TS;
  const float Wm1 = (float)a.W - 1.f, Hm1 = (float)a.H - 1.f;
  const float s_abs = sf[F_S];
  const int kc1 = half;                       // this thread's 16-column chunk of a 32-wide layer
  const int uq = v + V * half;                // pooling: this thread's first channel quad (stride 2 V)

  for (int tile = blockIdx.x * NG2 + grp; tile < ntiles; tile += gridDim.x * NG2) {
    const int p = tile * TS + sl;
    const bool active = lane_ok && (p < a.N);
    const int base = active ? row - v : (row & ~31);
    uint32_t* mvps = mvp + (active ? sl : (rm.packed ? (row >> 5) * rm.spw : 0)) * MVP;
    float4* sp = reinterpret_cast<float4*>(a.stash) + (size_t)tile * (ST_PLANES * GROUP) + row;
    const bool save = SAVE && active;

    // ---------------- projection, ray_diff (both halves), gather (own channels) ----------------
    float xh[XH];
    float rd[4];
    float mk = 0.f, ggx = 0.f, ggy = 0.f;
    if (active) {
      float X, Y, Z;
      load_point(a.pts, p, X, Y, Z);
      const ViewGeom g = view_geometry(X, Y, Z, s_cam + 16 * v, s_cam + 16 * V, Wm1, Hm1);
      if (half == 0) gather_half<0>(g, v, a.H, a.W, a.fh, a.fw, a.imgs, a.feat, xh);
      else gather_half<1>(g, v, a.H, a.W, a.fh, a.fw, a.imgs, a.feat, xh);
      rd[0] = g.rd[0]; rd[1] = g.rd[1]; rd[2] = g.rd[2]; rd[3] = g.rd[3];
      mk = g.mask; ggx = g.gx; ggy = g.gy;
    } else {
#pragma unroll
      for (int c = 0; c < XH; ++c) xh[c] = 0.f;
      rd[0] = rd[1] = rd[2] = rd[3] = 0.f;
    }
    const float rgb_in0 = xh[0], rgb_in1 = xh[1], rgb_in2 = xh[2];        // meaningful in half A

    // ---------------- ray_dir_fc.0 (K = 4) on the CUDA cores: 8 outputs per half -> 4 operand words ----------------
    {
      float a1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int o = 8 * half + j;
        a1[j] = sf[F_DIR0_B + o];
#pragma unroll
        for (int k = 0; k < 4; ++k) a1[j] = fmaf(rd[k], sf[F_DIR0_W + k * 16 + o], a1[j]);
        a1[j] = elu_fast(a1[j]);
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) pack_pair<NPASS>(a1[2 * j], a1[2 * j + 1], hi[j], lo[j]);
      tmem_st4(tl + C_A + 4 * half, hi);
      if (NPASS == 3) tmem_st4(tl + C_ALO + 4 * half, lo);
    }
    NFB_TC2_ISSUE(L_DIR2, 0, 1, false, 0);

    // ---------------- pooling weights (overlaps the MMA) ----------------
    float w, n_valid;
    {
      const float e = a.anti_alias ? (float)exp((double)__fmul_rn(s_abs, __fsub_rn(rd[3], 1.f))) : 1.f;
      if (half == 0) *reinterpret_cast<float2*>(ex + row * EXQ) = make_float2(e, mk);
      NFB_EX2_SYNC();
      float mn = 3.4e38f, nv = 0.f;
      for (int u = 0; u < V; ++u) {
        const float2 q = *reinterpret_cast<const float2*>(ex + (base + u) * EXQ);
        mn = fminf(mn, q.x);
        nv += q.y;
      }
      if (!a.anti_alias) mn = 0.f;
      float sum = 0.f;
      for (int u = 0; u < V; ++u) {
        const float2 q = *reinterpret_cast<const float2*>(ex + (base + u) * EXQ);
        sum += (q.x - mn) * q.y;
      }
      w = (e - mn) * mk / (sum + 1e-8f);
      n_valid = nv;
      NFB_EX2_SYNC();
    }

    // ---------------- x0 = rgb_feat + direction_feat (own channels) ----------------
    NFB_TC_WAIT();
    if (half == 0) {
      float df[16];
      epi16(tl, 0, sf + F_B_DIR2, df);
#pragma unroll
      for (int j = 0; j < 16; ++j) xh[j] += df[j];
      uint32_t r4[4];
      tmem_ld4u(tl + C_D + 16, r4);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 3; ++j) xh[16 + j] += elu_fast(__uint_as_float(r4[j]) + sf[F_B_DIR2 + 16 + j]);
    } else {
      float df[16];
      epi16(tl, 19, sf + F2_B_DIR2B, df);
#pragma unroll
      for (int j = 0; j < 16; ++j) xh[j] += df[j];
    }

    // stash + exchange rows: x0[0..19) from half A, x0[19..35) (+ w in slot 35) from half B
    if (half == 0) {
      if (save) {
#pragma unroll
        for (int j = 0; j < 4; ++j) __stcs(sp + (SP_X0 + j) * GROUP, make_float4(xh[4 * j], xh[4 * j + 1], xh[4 * j + 2], xh[4 * j + 3]));
        float* p4 = reinterpret_cast<float*>(sp + (SP_X0 + 4) * GROUP);
        __stcs(p4, xh[16]); __stcs(p4 + 1, xh[17]); __stcs(p4 + 2, xh[18]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(ex + row * EXQ + 4 * j) = make_float4(xh[4 * j], xh[4 * j + 1], xh[4 * j + 2], xh[4 * j + 3]);
      ex[row * EXQ + 16] = xh[16]; ex[row * EXQ + 17] = xh[17]; ex[row * EXQ + 18] = xh[18];
    } else {
      if (save) {
        __stcs(reinterpret_cast<float*>(sp + (SP_X0 + 4) * GROUP) + 3, xh[0]);
#pragma unroll
        for (int j = 0; j < 3; ++j) __stcs(sp + (SP_X0 + 5 + j) * GROUP, make_float4(xh[1 + 4 * j], xh[2 + 4 * j], xh[3 + 4 * j], xh[4 + 4 * j]));
        __stcs(sp + (SP_X0 + 8) * GROUP, make_float4(xh[13], xh[14], xh[15], 0.f));
      }
      ex[row * EXQ + 19] = xh[0];
#pragma unroll
      for (int j = 0; j < 3; ++j) *reinterpret_cast<float4*>(ex + row * EXQ + 20 + 4 * j) = make_float4(xh[1 + 4 * j], xh[2 + 4 * j], xh[3 + 4 * j], xh[4 + 4 * j]);
      *reinterpret_cast<float4*>(ex + row * EXQ + 32) = make_float4(xh[13], xh[14], xh[15], w);
    }
    NFB_EX2_SYNC();
    // ---------------- first pooling: weighted mean / variance over views (quads uq, uq + 2V, ...) ----------------
    if (active) {
      __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(mvps);
      __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(mvps + 36);
      pool4x<9, POOL_MEAN_VAR>(ex + base * EXQ, V, uq, 2 * V, 35, 1.f, [&](int q, const float4& m4, const float4& v4) {
        const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 4 * q + i;
          if (c < NFB_ROW_CH) {
            const __nv_bfloat16 mh = __float2bfloat16_rn(mm[i]), vh = __float2bfloat16_rn(vv[i]);
            ph[c] = mh;
            ph[35 + c] = vh;
            if (NPASS == 3) {
              pl[c] = __float2bfloat16_rn(mm[i] - __bfloat162float(mh));
              pl[35 + c] = __float2bfloat16_rn(vv[i] - __bfloat162float(vh));
            }
          }
        }
      });
    }
    NFB_EX2_SYNC();

    // ---------------- base_fc.0 : K' = [stats 0..69 | 0 0 | x[0..19) 0 | 0000 | x[19..35)] -> 64 ----------------
    // operand words: 0..34 stats (shared memory), 35 zero, 36..45 half A's x pairs, 46..47 zero, 48..55 half B's x pairs
    {
      uint32_t hi[16], lo[16];
      // words 16 * half .. + 16 of the statistics (k' < 64)
      {
        const uint4* s4 = reinterpret_cast<const uint4*>(mvps + 16 * half);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 t4 = s4[j];
          hi[4 * j] = t4.x; hi[4 * j + 1] = t4.y; hi[4 * j + 2] = t4.z; hi[4 * j + 3] = t4.w;
        }
        if (NPASS == 3) {
          const uint4* l4 = reinterpret_cast<const uint4*>(mvps + 36 + 16 * half);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 t4 = l4[j];
            lo[4 * j] = t4.x; lo[4 * j + 1] = t4.y; lo[4 * j + 2] = t4.z; lo[4 * j + 3] = t4.w;
          }
        }
      }
      tmem_st16w(tl + C_A + 16 * half, hi);
      if (NPASS == 3) {
        tmem_st16w(tl + C_ALO + 16 * half, lo);
        NFB_TC2_ISSUE(L_BASE0, 0, 4, false, 0);              // A region holds K = 64 per round
        NFB_TC_WAIT();
      }
      constexpr int C2 = (NPASS == 3) ? 0 : 32;              // column of operand word 32 in this round
      if (half == 0) {
        hi[0] = mvps[32]; hi[1] = mvps[33]; hi[2] = mvps[34]; hi[3] = 0u;
        lo[3] = 0u;
        if (NPASS == 3) { lo[0] = mvps[36 + 32]; lo[1] = mvps[36 + 33]; lo[2] = mvps[36 + 34]; }
#pragma unroll
        for (int j = 0; j < 10; ++j) pack_pair<NPASS>(xh[2 * j], xh[2 * j + 1], hi[4 + j], lo[4 + j]);    // xh[19] = 0 pad
        hi[14] = hi[15] = 0u;
        lo[14] = lo[15] = 0u;
        tmem_st16w(tl + C_A + C2, hi);
        if (NPASS == 3) tmem_st16w(tl + C_ALO + C2, lo);
      } else {
        uint32_t h8[8], l8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pack_pair<NPASS>(xh[2 * j], xh[2 * j + 1], h8[j], l8[j]);
        tmem_st8(tl + C_A + C2 + 16, h8);
        if (NPASS == 3) tmem_st8(tl + C_ALO + C2 + 16, l8);
      }
      if (NPASS == 3) NFB_TC2_ISSUE(L_BASE0, 4, 7, true, 1);
      else NFB_TC2_ISSUE(L_BASE0, 0, 7, false, 1);
      NFB_TC_WAIT();
    }

    // ---------------- base_fc.2 (64 -> 32): this half's 32 inputs = chunks 2 half, 2 half + 1 ----------------
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int kc = 2 * half + i;
      float h[16];
      uint32_t q[8];
      epi16_s<SAVE>(tl, 16 * kc, sf + F_B_BASE0 + 16 * kc, h, q);
      if (save) st_codes8(sp, SP_H1 + 2 * kc, q);
      a_store16<NPASS>(tl, kc, h);
    }
    NFB_TC2_ISSUE(L_BASE2, 0, 4, false, 2);
    NFB_TC_WAIT();

    // ---------------- vis_fc (32 -> 32 -> 33) on x1 * w: 16 columns per half ----------------
    float x1[16];
    {
      float h[16];
      uint32_t q[8];
      epi16_s<SAVE>(tl, 16 * kc1, sf + F_B_BASE2 + 16 * kc1, h, q);
      if (save) st_codes8(sp, SP_X1 + 2 * kc1, q);
      float t[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        x1[j] = h[j];
        t[j] = h[j] * w;
      }
      a_store16<NPASS>(tl, kc1, t);
    }
    NFB_TC2_ISSUE(L_VIS0, 0, 2, false, 1);
    NFB_TC_WAIT();
    {
      float h[16];
      uint32_t q[8];
      epi16_s<SAVE>(tl, 16 * kc1, sf + F_B_VIS0 + 16 * kc1, h, q);
      if (save) st_codes8(sp, SP_HV + 2 * kc1, q);
      a_store16<NPASS>(tl, kc1, h);
    }
    NFB_TC2_ISSUE(L_VIS2, 0, 2, false, 2);
    NFB_TC_WAIT();

    // x2 = x1 + x_res ; vis1 = sigmoid(xv[32]) * mask (both halves) ; vis_fc2 on x2 * vis1
    float vis1, sg1;
    uint32_t xvq16 = 0u;
    {
      uint32_t r4[4];
      tmem_ld4u(tl + C_D + 32, r4);
      tmem_ld_wait();
      const float2 y2 = elu_code2(make_float2(__uint_as_float(r4[0]) + sf[F_B_VIS2 + 32], 0.f), xvq16);
      sg1 = sigmoid_f(y2.x);
      vis1 = sg1 * mk;
    }
    {
      float h[16];
      uint32_t q[8];
      epi16_s<SAVE>(tl, 16 * kc1, sf + F_B_VIS2 + 16 * kc1, h, q);
      if (save) st_codes8(sp, SP_XV + 2 * kc1, q);
      float t[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        x1[j] += h[j];                      // x2
        t[j] = x1[j] * vis1;
      }
      if (save) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          __stcs(sp + (SP_X2 + 4 * kc1 + j) * GROUP, make_float4(x1[4 * j], x1[4 * j + 1], x1[4 * j + 2], x1[4 * j + 3]));
      }
      a_store16<NPASS>(tl, kc1, t);
    }
    NFB_TC2_ISSUE(L_VISB0, 0, 2, false, 3);
    NFB_TC_WAIT();
    float vis2, sg2;
    {
      float h[16];
      uint32_t q[8];
      epi16_s<SAVE>(tl, 16 * kc1, sf + F_B_VISB0 + 16 * kc1, h, q);
      if (save) st_codes8(sp, SP_HV2 + 2 * kc1, q);
      float zp = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) zp = fmaf(h[j], sf[F_W_VISB2 + 16 * kc1 + j], zp);
      // the other half of the dot product sits in the partner thread (warp q <-> q + 4): exchange through the (free) exchange row
      ex[row * EXQ + half] = zp;
      named_bar_sync(pair_id, 64);
      const float2 zz = *reinterpret_cast<const float2*>(ex + row * EXQ);
      sg2 = sigmoid_f(sf[F_B_VISB2] + zz.x + zz.y);
      vis2 = sg2 * mk;
    }

    // ---------------- rgb_fc on [x2, vis2, ray_diff] (37 -> 16 -> 8 -> 1) ----------------
    a_store16<NPASS>(tl, kc1, x1);
    if (half == 0) {
      float t[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) t[j] = 0.f;
      t[0] = vis2; t[1] = rd[0]; t[2] = rd[1]; t[3] = rd[2]; t[4] = rd[3];
      a_store16<NPASS>(tl, 2, t);
    }
    NFB_TC2_ISSUE(L_RGB0, 0, 3, false, 3);         // its group barrier also orders the z exchange above before the rows are rewritten
    // half B: publish its x2 columns for the second pooling while the MMA runs
    if (half == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(ex + row * EXQ + 16 + 4 * j) = make_float4(x1[4 * j], x1[4 * j + 1], x1[4 * j + 2], x1[4 * j + 3]);
    }
    NFB_TC_WAIT();
    if (half == 0) {
      float logit;
      float g1[16];
      uint32_t q1[8];
      epi16_s<SAVE>(tl, 0, sf + F_B_RGB0, g1, q1);
      float g2[8];
      load_bias<8>(g2, sf + F_B_RGB2);
      dense_acc<16, 8>(sf + F_W_RGB2, g1, g2);
      if (SAVE) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) elu_with_t(g2[j], g2[j], t[j]);
        if (save) {
          st_codes8(sp, SP_G1, q1);
          st_plane(sp, SP_G2, pack_t(t[0], t[1]), pack_t(t[2], t[3]), pack_t(t[4], t[5]), pack_t(t[6], t[7]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) g2[j] = elu_fast(g2[j]);
      }
      logit = dot_row<8>(g2, sf + F_W_RGB4) + sf[F_B_RGB4];
      if (mk == 0.f) logit = -1e9f;
      if (save) {
        __stcs(sp + SP_SA * GROUP, make_float4(w, mk, sg1, sg2));
        __stcs(sp + SP_SB * GROUP, make_float4(logit, ggx, ggy, rgb_in0));
        st_plane(sp, SP_SC, __float_as_uint(rgb_in1), __float_as_uint(rgb_in2), xvq16, 0u);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(ex + row * EXQ + 4 * j) = make_float4(x1[4 * j], x1[4 * j + 1], x1[4 * j + 2], x1[4 * j + 3]);
      *reinterpret_cast<float2*>(ex + row * EXQ + 32) = make_float2(vis2, logit);
      *reinterpret_cast<float4*>(side + row * SIDE) = make_float4(rgb_in0, rgb_in1, rgb_in2, 0.f);
    }
    NFB_EX2_SYNC();

    // ---------------- second pooling, blending, output ----------------
    if (active) {
      float Dn = 1e-8f;
      for (int u = 0; u < V; ++u) Dn += ex[(base + u) * EXQ + 32];
      const float invD = 1.f / Dn;
      float* out = a.ps + (size_t)p * NFB_PS_STRIDE;
      pool4x<8, POOL_MEAN_VAR>(ex + base * EXQ, V, uq, 2 * V, 32, invD, [&](int q, const float4& m4, const float4& v4) {
        *reinterpret_cast<float4*>(out + PS_MEAN + 4 * q) = m4;
        *reinterpret_cast<float4*>(out + PS_VAR + 4 * q) = v4;
      });
      if (v == 0 && half == 1) {            // half B has the lighter tail
        float mx = -3.4e38f;
        for (int u = 0; u < V; ++u) mx = fmaxf(mx, ex[(base + u) * EXQ + 33]);
        float se = 0.f;
        for (int u = 0; u < V; ++u) se += __expf(ex[(base + u) * EXQ + 33] - mx);
        const float inv_se = 1.f / se;
        float wsum = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f;
        for (int u = 0; u < V; ++u) {
          const float2 vl = *reinterpret_cast<const float2*>(ex + (base + u) * EXQ + 32);
          const float4 c4 = *reinterpret_cast<const float4*>(side + (base + u) * SIDE);
          wsum += vl.x * invD;
          const float b = __expf(vl.y - mx) * inv_se;
          r0 = fmaf(b, c4.x, r0);
          r1 = fmaf(b, c4.y, r1);
          r2 = fmaf(b, c4.z, r2);
        }
        *reinterpret_cast<float4*>(out + 64) = make_float4(wsum / (float)V, r0, r1, r2);   // PS_WMEAN, PS_RGB
        *reinterpret_cast<float4*>(out + 68) = make_float4(n_valid, 0.f, 0.f, 0.f);        // PS_NVALID
      }
    }
    NFB_EX2_SYNC();                          // exchange buffer is reused by the next tile
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*s_tmem, 512);
}

template <int NPASS, bool SAVE>
int launch_view_tc_fwd2(const ViewArgs& a, cudaStream_t st) {
  constexpr size_t smem = smem_bytes2<NPASS>();
  cudaError_t e = cudaFuncSetAttribute(k_view_tc_fwd2<NPASS, SAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "k_view_tc_fwd2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int TS = row_map(a.V).TS;
  const int ntiles = (a.N + TS - 1) / TS;
  int grid = (ntiles + NG2 - 1) / NG2;
  const int cap = nfb_num_sms();
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  k_view_tc_fwd2<NPASS, SAVE><<<grid, GT * NG2, smem, st>>>(a);
  NFB_CHECK_LAUNCH("k_view_tc_fwd2");
  return NFB_OK;
}

}  // namespace nfbvtc2

// defined in nfb_view_tc2_inst.cu (one instantiation per translation unit)
int nfb_launch_view_tc_fwd2_p1(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_tc_fwd2_p3(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_tc_fwd2_p1_save(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_tc_fwd2_p3_save(const nfbview::ViewArgs& a, cudaStream_t st);
