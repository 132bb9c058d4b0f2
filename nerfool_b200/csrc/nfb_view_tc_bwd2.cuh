// IBRNet view stage on tcgen05: data-gradient FROM THE ACTIVATION STASH (fused mode).
//
// The forward kernel (nfb_view_tc.cuh, SAVE) wrote, per (sample, view) row, what the backward needs: x0, x2, a
// handful of scalars and the ELU-derivative codes of every hidden layer (768 B per row, 16-byte vectors laid out so
// that neighbouring rows are neighbours in memory).  Reading that back costs 768 B/row of HBM traffic; recomputing it
// (nfb_view_tc_bwd.cuh) costs the whole forward again -- ~6 k instructions per row, 8 warps per SM because the
// recompute needs 255 registers and 256 TMEM columns per row group.  On a B200 (6.5 TB/s measured) the stash wins:
// this kernel is a pure backward, 128 registers, 4 row groups (16 warps) per SM.
//
// Same row mapping as the forward (one thread per row, 128-row groups = TS samples x V views).  The backward dense
// layers dX = dY W read the forward weight tiles MN-major (instruction-descriptor bit 16).
// TMEM columns of a group (128): D [0,64) | A hi [64,96) | A lo [96,128).  base_fc.0's backward (112 outputs) is
// issued as two MMAs of 64 + 48 output columns from the same A operand.
#pragma once
#include "nfb_view_tc.cuh"

#ifndef NFB_VTC_COOP_SCATTER
#define NFB_VTC_COOP_SCATTER 1     // quarter-warp cooperative scatter (scatter_row_coop, nfb_geom.cuh); 0 = the pairwise de-duplicated per-lane scatter: 123.0 vs 110.4 ms
#endif
namespace nfbvtcs {
using namespace nfbtc;
using namespace nfbvtc;
using nfbview::ViewArgs;

constexpr int DPS = 100;     // staged cotangent row: d_mean[32] d_var[32] d_wmean d_rgb[3] | forward mean[32]
                             // (stride 100 = 4 mod 32 words: rows of the 8 samples a warp touches hit distinct banks)

template <int NPASS>
__host__ __device__ constexpr size_t smem_bytes_bwd2() {
  return (size_t)B_SET_BYTES * (NPASS == 3 ? 2 : 1) +
         sizeof(float) * (F_TOTAL + NG * GROUP * EXQ + NG * TS_MAX * MVS + NG * TS_MAX * DPS) + NG * 8 + 16;
}

__device__ __forceinline__ void d_raw16(uint32_t tl, int col, float (&y)[16]) {
  tmem_ld16(tl + C_D + col, y);
  tmem_ld_wait();
}

// dX[128][N0 .. N0 + NW) = dY[128][layer_n] W[layer_n][N0 .. N0 + NW): MMA N = NW, MMA K = layer_n; the B operand is
// the forward tile read MN-major (core matrices 128 B apart along MMA-K, layer_n * 16 B apart along MMA-N)
template <int NPASS, int LAYER, int N0, int NW>
__device__ __forceinline__ void issue_bwd(uint32_t tb, uint32_t sB_addr) {
  constexpr int NO = layer_n(LAYER);
  constexpr uint32_t idesc = idesc_bf16(128, NW) | (1u << 16);
  const uint32_t bhi = sB_addr + layer_off(LAYER) + (N0 / 8) * (NO * 16), blo = bhi + B_SET_BYTES;
#pragma unroll
  for (int ks = 0; ks < NO / 16; ++ks) {
    const uint64_t dh = smem_desc(bhi + ks * 256, 128, NO * 16);
    const uint32_t ah = tb + C_A + 8 * ks;
    mma_ts(tb + C_D, ah, dh, idesc, ks > 0);
    if (NPASS == 3) {
      const uint64_t dl = smem_desc(blo + ks * 256, 128, NO * 16);
      mma_ts(tb + C_D, tb + C_ALO + 8 * ks, dh, idesc, true);
      mma_ts(tb + C_D, ah, dl, idesc, true);
    }
  }
}

#define NFB_TCS_BWD(LAYER, N0, NW, ISSUER)                                    \
  do {                                                                        \
    tmem_st_wait();                                                           \
    fence_before_sync();                                                      \
    named_bar_sync(bar_id, GROUP);                                            \
    if (tg == 32 * (ISSUER)) {      /* issuer rotates over the group's warps: see NFB_TC_ISSUE */ \
      fence_after_sync();                                                     \
      issue_bwd<NPASS, LAYER, N0, NW>(tb, sB_addr);                           \
      mma_commit(mbar);                                                       \
    }                                                                         \
  } while (0)
#define NFB_TCS_WAIT()         \
  do {                         \
    mbar_wait(mbar, phase);    \
    phase ^= 1u;               \
    fence_after_sync();        \
  } while (0)

__device__ __forceinline__ void ld_codes8(const float4* sp, int plane, uint32_t (&q)[8]) {
  const float4 a = __ldcs(sp + plane * GROUP), b = __ldcs(sp + (plane + 1) * GROUP);
  q[0] = __float_as_uint(a.x); q[1] = __float_as_uint(a.y); q[2] = __float_as_uint(a.z); q[3] = __float_as_uint(a.w);
  q[4] = __float_as_uint(b.x); q[5] = __float_as_uint(b.y); q[6] = __float_as_uint(b.z); q[7] = __float_as_uint(b.w);
}

__device__ __forceinline__ void ld_codes16(const float4* sp, int plane, uint32_t (&q)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 a = __ldcs(sp + (plane + i) * GROUP);
    q[4 * i] = __float_as_uint(a.x); q[4 * i + 1] = __float_as_uint(a.y);
    q[4 * i + 2] = __float_as_uint(a.z); q[4 * i + 3] = __float_as_uint(a.w);
  }
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int NPASS>
__global__ void __launch_bounds__(GROUP * NG, 1) k_view_tc_bwd_stash(ViewArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sB = smem_raw;
  float* sf = reinterpret_cast<float*>(smem_raw + (size_t)B_SET_BYTES * (NPASS == 3 ? 2 : 1));
  float* ex_all = sf + F_TOTAL;
  float* mv_all = ex_all + NG * GROUP * EXQ;
  float* dp_all = mv_all + NG * TS_MAX * MVS;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(dp_all + NG * TS_MAX * DPS);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + NG);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int grp = tid / GROUP, tg = tid % GROUP;
  float* ex = ex_all + (size_t)grp * GROUP * EXQ;
  float* mv = mv_all + (size_t)grp * TS_MAX * MVS;
  float* dpb = dp_all + (size_t)grp * TS_MAX * DPS;
  const int bar_id = 1 + grp;
  uint64_t* mbar = s_bar + grp;

  if (warp == 0) tmem_alloc(s_tmem, TMEM_ALLOC);
  if (tid == 0) {
    for (int g = 0; g < NG; ++g) mbar_init(s_bar + g, 1);
    mbar_init_fence();
  }
  {
    const float* p = a.params;
    load_tile<NPASS>(sB, L_DIR2, p + P_DIR2_W, 35, 16, tid, blockDim.x);      // keeps the tile offsets of the forward
    load_tile<NPASS>(sB, L_BASE0, p + P_BASE0_W, 64, 105, tid, blockDim.x);
    load_tile<NPASS>(sB, L_BASE2, p + P_BASE2_W, 32, 64, tid, blockDim.x);
    load_tile<NPASS>(sB, L_VIS0, p + P_VIS0_W, 32, 32, tid, blockDim.x);
    load_tile<NPASS>(sB, L_VIS2, p + P_VIS2_W, 33, 32, tid, blockDim.x);
    load_tile<NPASS>(sB, L_VISB0, p + P_VISB0_W, 32, 32, tid, blockDim.x);
    load_tile<NPASS>(sB, L_RGB0, p + P_RGB0_W, 16, 37, tid, blockDim.x);
    load_side_tables(sf, p, tid, blockDim.x);
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
  const RowMap rm = row_map(V);                // the forward's tile mapping (the stash is indexed by tile and thread)
  const bool warp_local = rm.packed;           // NFB_EX_SYNC (nfb_view_tc.cuh): warp barrier when a sample's rows share a warp
  const int TS = rm.TS;
  int sl, v, partner = tid & 31;
  bool lane_ok, pair_leader = true;
  if (rm.packed) {
    const int wq = tg >> 5, l = tg & 31, si = l / V;
    v = l - si * V; sl = wq * rm.spw + si; lane_ok = si < rm.spw;
    if (lane_ok) {                             // scatter pairs: samples (2k, 2k + 1) of the warp, same view
      pair_leader = (si & 1) == 0;
      if (!pair_leader) partner = l - V;
      else if (si + 1 < rm.spw) partner = l + V;
    }
  } else {
    sl = tg / V; v = tg - sl * V; lane_ok = sl < TS;
  }
  const int ntiles = (a.N + TS - 1) / TS;

  for (int tile = blockIdx.x * NG + grp; tile < ntiles; tile += gridDim.x * NG) {
    const int p = tile * TS + sl;
    const bool active = lane_ok && (p < a.N);
    // rows without a sample read rows / slots of their OWN warp (see the forward kernel): covered by the warp-level fences
    const int base = active ? tg - v : (tg & ~31);
    const int sl_safe = active ? sl : (rm.packed ? (tg >> 5) * rm.spw : 0);
    float* mvs = mv + sl_safe * MVS;
    const float* dp = dpb + sl_safe * DPS;
    const float4* sp = reinterpret_cast<const float4*>(a.stash) + (size_t)tile * (ST_PLANES * GROUP) + tg;

    // ---------------- stage the cotangents / forward means of the tile's samples (asynchronous copies) ----------------
    {
      const int p0 = tile * TS;
      const int ns = (a.N - p0 < TS) ? (a.N - p0) : TS;
      for (int i = tg; i < ns * 25; i += GROUP) {        // 17 float4 of d_ps (68 floats) + 8 float4 of ps (mean)
        const int s = i / 25, j = i - s * 25;
        const float4* src = (j < 17) ? reinterpret_cast<const float4*>(a.d_ps + (size_t)(p0 + s) * NFB_PS_STRIDE) + j
                                     : reinterpret_cast<const float4*>(a.ps + (size_t)(p0 + s) * NFB_PS_STRIDE) + (j - 17);
        cp_async16(dpb + s * DPS + 4 * j, src);
      }
      // pull the NEXT tile of this group towards L2 while this one is processed: its stash planes (one line per 8 rows)
      // and its cotangent rows
      const int nt = tile + gridDim.x * NG;
      if (nt < ntiles) {
        if ((tg & 7) == 0) {
          const float4* np = reinterpret_cast<const float4*>(a.stash) + (size_t)nt * (ST_PLANES * GROUP) + tg;
#pragma unroll 5
          for (int pl = 0; pl < SP_H1; ++pl) prefetch_l2(np + pl * GROUP);          // x0, x2, scalars
          prefetch_l2(np + SP_G1 * GROUP);
          prefetch_l2(np + (SP_G1 + 1) * GROUP);
          prefetch_l2(np + SP_G2 * GROUP);
        }
        const int q0 = nt * TS;
        const int nn = (a.N - q0 < TS) ? (a.N - q0) : TS;
        for (int i = tg; i < nn * 5; i += GROUP) {       // 288-byte rows: 3 lines of d_ps, 2 of ps (first 128 B = means)
          const int s = i / 5, j = i - s * 5;
          prefetch_l2(j < 3 ? reinterpret_cast<const char*>(a.d_ps + (size_t)(q0 + s) * NFB_PS_STRIDE) + 128 * j
                            : reinterpret_cast<const char*>(a.ps + (size_t)(q0 + s) * NFB_PS_STRIDE) + 128 * (j - 3));
        }
      }
    }

    // ---------------- per-row scalars ----------------
    float w = 0.f, mk = 0.f, sg1 = 0.f, sg2 = 0.f, logit = 0.f, gx = 0.f, gy = 0.f, rgb_in0 = 0.f, rgb_in1 = 0.f, rgb_in2 = 0.f;
    uint32_t xvq16 = 0x00008000u;   // code of ELU' = 1
    if (active) {
      const float4 sa = __ldcs(sp + SP_SA * GROUP), sb = __ldcs(sp + SP_SB * GROUP), sc = __ldcs(sp + SP_SC * GROUP);
      w = sa.x; mk = sa.y; sg1 = sa.z; sg2 = sa.w;
      logit = sb.x; gx = sb.y; gy = sb.z; rgb_in0 = sb.w;
      rgb_in1 = sc.x; rgb_in2 = sc.y; xvq16 = __float_as_uint(sc.z);
    }
    const float vis1 = sg1 * mk, vis2 = sg2 * mk;

    // ---------------- mean0 of the first pooling (needed by its backward): exchange x0 and w ----------------
    {
      if (active) {
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          float4 q = __ldcs(sp + (SP_X0 + j) * GROUP);
          if (j == 8) q.w = w;                              // slot 35: the pooling weight
          *reinterpret_cast<float4*>(ex + tg * EXQ + 4 * j) = q;
        }
      } else {
        ex[tg * EXQ + 35] = 0.f;
      }
      NFB_EX_SYNC();
    }
    float wsum0 = 0.f;
    for (int u = 0; u < V; ++u) wsum0 += ex[(base + u) * EXQ + 35];
    if (active) {
      pool4<9, POOL_MEAN>(ex + base * EXQ, V, v, 35, 1.f, [&](int q, const float4& m4, const float4&) {
        *reinterpret_cast<float4*>(mvs + 4 * q) = m4;       // mvs[35] (quad 8, lane 3) is a pad slot
      });
    }
    NFB_EX_SYNC();

    // ---------------- exchange vis2 / logit / rgb_in ; blending softmax ----------------
    // slots 0..4 of the row: vis2, logit, rgb_in[3]  (the x0 rows above are dead: mean0 sits in mvs)
    *reinterpret_cast<float4*>(ex + tg * EXQ) = make_float4(vis2, logit, rgb_in0, rgb_in1);
    ex[tg * EXQ + 4] = rgb_in2;
    cp_async_wait_all();                    // the staged cotangent rows are read after this barrier
    named_bar_sync(bar_id, GROUP);
    float Dsum = 1e-8f, mx = -3.4e38f;
    for (int u = 0; u < V; ++u) {
      const float2 q = *reinterpret_cast<const float2*>(ex + (base + u) * EXQ);
      Dsum += q.x;
      mx = fmaxf(mx, q.y);
    }
    const float invD = 1.f / Dsum;
    float se = 0.f, w2sum = 0.f;
    for (int u = 0; u < V; ++u) {
      const float2 q = *reinterpret_cast<const float2*>(ex + (base + u) * EXQ);
      se += __expf(q.y - mx);
      w2sum += q.x * invD;
    }
    const float inv_se = 1.f / se;
    const float w2 = vis2 * invD;
    const float d_r0 = dp[65], d_r1 = dp[66], d_r2 = dp[67];
    const float d_wmean = dp[64];

    // (1) blending softmax
    const float blend = __expf(logit - mx) * inv_se;
    float d_logit;
    {
      float bt = 0.f;
      for (int u = 0; u < V; ++u) {
        const float4 q = *reinterpret_cast<const float4*>(ex + (base + u) * EXQ);
        const float b = __expf(q.y - mx) * inv_se;
        const float tu = q.z * d_r0 + q.w * d_r1 + ex[(base + u) * EXQ + 4] * d_r2;
        bt = fmaf(b, tu, bt);
      }
      const float tv = rgb_in0 * d_r0 + rgb_in1 * d_r1 + rgb_in2 * d_r2;
      d_logit = (mk != 0.f) ? blend * (tv - bt) : 0.f;
    }

    // (2) rgb_fc backward: the two small layers on the CUDA cores, rgb_fc.0 as MMA (48 outputs: d[x2 | vis2 | ray_diff])
    {
      uint32_t q1[8];
      ld_codes8(sp, SP_G1, q1);
      const float4 q2 = __ldcs(sp + SP_G2 * GROUP);
      const uint32_t g2q[4] = {__float_as_uint(q2.x), __float_as_uint(q2.y), __float_as_uint(q2.z), __float_as_uint(q2.w)};
      float dg2[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dg2[2 * j] = d_logit * sf[F_W_RGB4 + 2 * j] * elu_stash_lo(g2q[j]);
        dg2[2 * j + 1] = d_logit * sf[F_W_RGB4 + 2 * j + 1] * elu_stash_hi(g2q[j]);
      }
      float dg1[16];
      dense_T<16, 8>(sf + F_W_RGB2, dg2, dg1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mul_stash2(dg1[2 * j], dg1[2 * j + 1], q1[j]);
      }
      a_store16<NPASS>(tl, 0, dg1);
    }
    NFB_TCS_BWD(L_RGB0, 0, 48, 0);

    // (3) second pooling backward (overlaps the MMA)
    float x2[32];
    float d_x2[32];
    float d_w2 = d_wmean / (float)V;
    {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 q = __ldcs(sp + (SP_X2 + j) * GROUP);
        x2[4 * j] = q.x; x2[4 * j + 1] = q.y; x2[4 * j + 2] = q.z; x2[4 * j + 3] = q.w;
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
    NFB_EX_SYNC();          // all reads of slots 0..4 above are done
    ex[tg * EXQ + 1] = d_w2 * vis2;
    NFB_EX_SYNC();
    float d_vis2;
    {
      float sdv = 0.f;
      for (int u = 0; u < V; ++u) sdv += ex[(base + u) * EXQ + 1];
      d_vis2 = d_w2 * invD - sdv * invD * invD;
    }
    uint32_t cq[16];                        // ELU' codes of the next layer, loaded ahead of the MMA wait
    ld_codes16(sp, SP_HV2, cq);
    NFB_TCS_WAIT();
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
          const float2 wv = *reinterpret_cast<const float2*>(sf + F_W_VISB2 + 16 * kc + 2 * j);
          const float2 r = __fmul2_rn(__fmul2_rn(make_float2(dz, dz), wv), elu_stash2(cq[8 * kc + j]));
          dh[2 * j] = r.x; dh[2 * j + 1] = r.y;
        }
        a_store16<NPASS>(tl, kc, dh);
      }
      ld_codes16(sp, SP_XV, cq);            // stash loads of the NEXT layer issued before the group barrier: in flight while it waits
      NFB_TCS_BWD(L_VISB0, 0, 32, 1);
      NFB_TCS_WAIT();
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
          const float2 r = __fmul2_rn(make_float2(d_x2[16 * kc + 2 * j], d_x2[16 * kc + 2 * j + 1]), elu_stash2(cq[8 * kc + j]));
          dxv[2 * j] = r.x; dxv[2 * j + 1] = r.y;
        }
        a_store16<NPASS>(tl, kc, dxv);
      }
      float dxv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) dxv[j] = 0.f;
      dxv[0] = d_vis1 * mk * sg1 * (1.f - sg1) * elu_stash_lo(xvq16);
      a_store16<NPASS>(tl, 2, dxv);
    }
    ld_codes16(sp, SP_HV, cq);
    NFB_TCS_BWD(L_VIS2, 0, 32, 2);
    NFB_TCS_WAIT();
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      float dh[16];
      d_raw16(tl, 16 * kc, dh);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mul_stash2(dh[2 * j], dh[2 * j + 1], cq[8 * kc + j]);
      }
      a_store16<NPASS>(tl, kc, dh);
    }
    ld_codes16(sp, SP_X1, cq);
    NFB_TCS_BWD(L_VIS0, 0, 32, 3);
    NFB_TCS_WAIT();

    // (6) base_fc backward
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      float dt[16];
      d_raw16(tl, 16 * kc, dt);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 r = __fmul2_rn(__ffma2_rn(make_float2(dt[2 * j], dt[2 * j + 1]), make_float2(w, w),
                                               make_float2(d_x2[16 * kc + 2 * j], d_x2[16 * kc + 2 * j + 1])),
                                    elu_stash2(cq[8 * kc + j]));
        dt[2 * j] = r.x; dt[2 * j + 1] = r.y;
      }
      a_store16<NPASS>(tl, kc, dt);
    }
    NFB_TCS_BWD(L_BASE2, 0, 64, 0);
    uint32_t hq[32];
    ld_codes16(sp, SP_H1, cq);
    {
      uint32_t t16[16];
      ld_codes16(sp, SP_H1 + 4, t16);
#pragma unroll
      for (int j = 0; j < 16; ++j) { hq[j] = cq[j]; hq[16 + j] = t16[j]; }
    }
    NFB_TCS_WAIT();
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      float dh[16];
      d_raw16(tl, 16 * kc, dh);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mul_stash2(dh[2 * j], dh[2 * j + 1], hq[8 * kc + j]);
      }
      a_store16<NPASS>(tl, kc, dh);
    }
    // base_fc.0 inputs [mean0 (35) | var0 (35) | x0 (35) | pad]: first MMA = input columns [0,64)
    NFB_TCS_BWD(L_BASE0, 0, 64, 1);
    NFB_TCS_WAIT();

    // (7) first pooling backward.  With Dm_c = sum_v d mean0_vc, Dv_c = sum_v d var0_vc:
    //   d x0_vc = dx_vc + w_v A_c + w_v x0_vc B_c,  B_c = 2 Dv_c,  A_c = Dm_c - 2 Dv_c mean0_c (2 - wsum)
    // d mean0 = columns 0..34 and d var0[0..29) = columns 35..63 of the first MMA; the rest comes from the second
    float dvar_lo[29];
    {
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        float t[16];
        d_raw16(tl, c0, t);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (c0 + j < 35) ex[tg * EXQ + c0 + j] = t[j];
          else dvar_lo[c0 + j - 35] = t[j];
        }
      }
    }
    // second MMA: input columns [64,112) -> D columns [0,48): d var0[29..35) then d x0[0..35)
    NFB_TCS_BWD(L_BASE0, 64, 48, 2);            // (its barrier also publishes the d mean0 rows written above)
    // exchange round 1 (overlaps the MMA): per-sample sums of d mean0, parked in the (now dead) cotangent staging row
    float* dpw = dpb + sl_safe * DPS;
    if (active) {
      pool4<9, POOL_SUM>(ex + base * EXQ, V, v, 0, 1.f, [&](int q, const float4& s4, const float4&) {
        *reinterpret_cast<float4*>(dpw + 4 * q) = s4;       // dpw[35] is a pad slot (d_var[3] of the dead cotangent row)
      });
    }
    NFB_EX_SYNC();
    float x0[36];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const float4 q = active ? __ldcs(sp + (SP_X0 + j) * GROUP) : make_float4(0.f, 0.f, 0.f, 0.f);
      x0[4 * j] = q.x; x0[4 * j + 1] = q.y; x0[4 * j + 2] = q.z; x0[4 * j + 3] = q.w;
    }
    NFB_TCS_WAIT();
    float d_row[NFB_ROW_CH];
    {
      // d var0: 29 values kept from the first MMA + 6 from the second
#pragma unroll
      for (int c = 0; c < 29; ++c) ex[tg * EXQ + c] = dvar_lo[c];
      float t[16];
      d_raw16(tl, 0, t);
#pragma unroll
      for (int j = 0; j < 6; ++j) ex[tg * EXQ + 29 + j] = t[j];
#pragma unroll
      for (int j = 6; j < 16; ++j) d_row[j - 6] = t[j];
      d_raw16(tl, 16, t);
#pragma unroll
      for (int j = 0; j < 16; ++j) d_row[10 + j] = t[j];
      d_raw16(tl, 32, t);
#pragma unroll
      for (int j = 0; j < 9; ++j) d_row[26 + j] = t[j];
    }
    NFB_EX_SYNC();
    if (active) {
      pool4<9, POOL_SUM>(ex + base * EXQ, V, v, 0, 1.f, [&](int q, const float4& s4, const float4&) {
        const float4 m0 = *reinterpret_cast<const float4*>(mvs + 4 * q), dm = *reinterpret_cast<const float4*>(dpw + 4 * q);
        const float k2 = 2.f - wsum0;
        *reinterpret_cast<float4*>(mvs + 4 * q) = make_float4(dm.x - 2.f * s4.x * m0.x * k2, dm.y - 2.f * s4.y * m0.y * k2,
                                                             dm.z - 2.f * s4.z * m0.z * k2, dm.w - 2.f * s4.w * m0.w * k2);
        *reinterpret_cast<float4*>(mvs + 36 + 4 * q) = make_float4(2.f * s4.x, 2.f * s4.y, 2.f * s4.z, 2.f * s4.w);
      });
    }
    NFB_EX_SYNC();
    if (active) {
#pragma unroll
      for (int c = 0; c < NFB_ROW_CH; ++c) d_row[c] += w * (mvs[c] + x0[c] * mvs[36 + c]);
      d_row[0] = fmaf(blend, d_r0, d_row[0]);
      d_row[1] = fmaf(blend, d_r1, d_row[1]);
      d_row[2] = fmaf(blend, d_r2, d_row[2]);
    }
    // (8) scatter (grid_sampler_2d backward w.r.t. the input)
#if NFB_VTC_COOP_SCATTER
    // the exchange rows are free here (fence above): they stage the hand-over of the cotangent rows to the quarter-warps
    scatter_row_coop(active, gx, gy, v, a.H, a.W, a.fh, a.fw, d_row, a.d_feat, a.d_imgs, ex, EXQ, tg);
#else
    if (warp_local && rm.spw >= 2) {
      // sample pairs (2k, 2k + 1) of a warp, same view: same texel quad -> one set of atomics for both
      scatter_row_paired(active, gx, gy, v, partner, pair_leader, a.H, a.W, a.fh, a.fw, d_row, a.d_feat, a.d_imgs);
    } else if (active) {
      ViewGeom g;
      g.gx = gx; g.gy = gy;
      scatter_row(g, v, a.H, a.W, a.fh, a.fw, d_row, a.d_feat, a.d_imgs);
    }
#endif
    named_bar_sync(bar_id, GROUP);          // exchange / statistics / staging buffers are reused by the next tile
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*s_tmem, TMEM_ALLOC);
}

template <int NPASS>
int launch_view_tc_bwd_stash(const ViewArgs& a, cudaStream_t st) {
  constexpr size_t smem = smem_bytes_bwd2<NPASS>();
  cudaError_t e = cudaFuncSetAttribute(k_view_tc_bwd_stash<NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "k_view_tc_bwd_stash: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int TS = row_map(a.V).TS;
  const int ntiles = (a.N + TS - 1) / TS;
  int grid = (ntiles + NG - 1) / NG;
  const int cap = nfb_num_sms();
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  k_view_tc_bwd_stash<NPASS><<<grid, GROUP * NG, smem, st>>>(a);
  NFB_CHECK_LAUNCH("k_view_tc_bwd_stash");
  return NFB_OK;
}

}  // namespace nfbvtcs

int nfb_launch_view_tc_bwd_stash_p1(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_tc_bwd_stash_p3(const nfbview::ViewArgs& a, cudaStream_t st);
