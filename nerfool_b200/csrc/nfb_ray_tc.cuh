// IBRNet ray stage on the 5th-generation tensor cores (tcgen05 + TMEM), forward and data-gradient:
// geometry_fc, + pos_encoding, 4-head d_k = 4 self-attention over the samples of a ray (row-masked), fc + residual +
// LayerNorm(eps 1e-6), sigma head (mlp_network.py:259-265, 69-119, 23-43).
//
// Mapping: one thread per sample; a 128-thread group owns floor(128 / S) whole rays per tile (S <= 128); rays of
// 129..256 samples are owned by a pair of groups sharing one 256-row attention buffer.  Dense layers run as 128 x N x K tcgen05.mma tiles exactly like the view
// stage (A operand = the rows' activations written to TMEM by their threads, B = weight tiles resident in shared
// memory, D read back with tcgen05.ld); the backward products dX = dY W read the same tiles MN-major.  The attention
// itself (d_k = 4 per head: too thin for an MMA) stays on the CUDA cores with K / V (backward: Q, dO, softmax
// statistics) of the group's rays in shared memory.
//
// TMEM columns of a group: D [0,64) | A hi [64,96) | A lo [96,128) | backward only: ELU'(h64) codes [128,160).
#pragma once
#include "nfb_dense.cuh"
#include "nfb_tc.cuh"

namespace nfbrtc {
using namespace nfbtc;

constexpr int GROUP = 128;
constexpr int RC_D = 0, RC_A = 64, RC_ALO = 96, RC_HQ = 128;

enum : int { RL_GEO0 = 0, RL_GEO2, RL_QKV, RL_FC, RL_OG0, RL_COUNT };
__host__ __device__ constexpr int rl_n(int l) { return l == RL_GEO0 ? 64 : l == RL_QKV ? 48 : 16; }
__host__ __device__ constexpr int rl_k(int l) { return (l == RL_GEO0 || l == RL_GEO2) ? 64 : 16; }
__host__ __device__ constexpr int rl_off(int l) {
  int o = 0;
  for (int i = 0; i < l; ++i) o += rl_n(i) * rl_k(i) * 2;
  return o;
}
constexpr int R_SET_BYTES = rl_off(RL_COUNT);   // 12800
static_assert(R_SET_BYTES % 16 == 0, "tile alignment");

enum : int {
  RF_B_GEO0 = 0,               // 64
  RF_WCOL = RF_B_GEO0 + 64,    // 64: geometry_fc.0.weight[:, 64] (the mean-weight input, handled on the CUDA cores)
  RF_B_GEO2 = RF_WCOL + 64,    // 16
  RF_LNW = RF_B_GEO2 + 16,     // 16
  RF_LNB = RF_LNW + 16,        // 16
  RF_B_OG0 = RF_LNB + 16,      // 16
  RF_W_OG2 = RF_B_OG0 + 16,    // 16
  RF_B_OG2 = RF_W_OG2 + 16,    // 4
  RF_TOTAL = RF_B_OG2 + 4
};

constexpr float INV_TEMP = 0.5f;     // 1 / sqrt(d_k), d_k = 4 (mlp_network.py:84)
constexpr float LN_EPS = 1e-6f;      // mlp_network.py:87
constexpr float LOG2E = 1.4426950408889634f;

template <int NPASS, bool BWD>
struct Cfg {
  static constexpr int NG = BWD ? 2 : 4;
  static constexpr int GC = BWD ? 256 : 128;
  // per-group shared memory (floats): K, V [128][16]; backward: Q, dO [128][16] and statistics [128][4] float4;
  // forward: Q [128][16], the pair exchange [128][24] and the per-warp key norms [4][4] of the two-query attention
  static constexpr int GROUP_FLOATS = BWD ? (4 * GROUP * 16 + GROUP * 16) : (3 * GROUP * 16 + GROUP * 24 + 16);
  static size_t smem(int S) {
    return (size_t)R_SET_BYTES * (NPASS == 3 ? 2 : 1) + sizeof(float) * (RF_TOTAL + (size_t)S * 16 + (size_t)NG * GROUP_FLOATS) +
           NG * 8 + 16;
  }
};

template <int NPASS>
static __device__ void load_rtile(uint8_t* sB, int layer, const float* __restrict__ w, int n_real, int row_stride, int tid, int nt) {
  const int N = rl_n(layer), K = rl_k(layer);
  uint8_t* hi = sB + rl_off(layer);
  uint8_t* lo = hi + R_SET_BYTES;
  for (int i = tid; i < N * K; i += nt) {
    const int n = i / K, k = i - n * K;
    const float v = (n < n_real) ? __ldg(w + n * row_stride + k) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const uint32_t off = (uint32_t)((k >> 3) * (N * 16) + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2);
    *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
    if (NPASS == 3) *reinterpret_cast<__nv_bfloat16*>(lo + off) = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

template <int NPASS>
__device__ __forceinline__ void r_put16(uint32_t tl, int kc, const float (&v)[16]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (NPASS == 3) split_bf16(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
    else hi[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
  }
  tmem_st8(tl + RC_A + 8 * kc, hi);
  if (NPASS == 3) tmem_st8(tl + RC_ALO + 8 * kc, lo);
}
__device__ __forceinline__ void r_ld16(uint32_t tl, int col, float (&y)[16]) {
  tmem_ld16(tl + RC_D + col, y);
  tmem_ld_wait();
}

// forward: D[128][N] = A[128][K] W^T ; backward: D[128][K] = A[128][N] W (MN-major read of the same tile)
template <int NPASS, int LAYER>
__device__ __forceinline__ void r_issue_fwd(uint32_t tb, uint32_t sB_addr) {
  constexpr int N = rl_n(LAYER), K = rl_k(LAYER);
  constexpr uint32_t idesc = idesc_bf16(128, N);
  const uint32_t bhi = sB_addr + rl_off(LAYER), blo = bhi + R_SET_BYTES;
#pragma unroll
  for (int ks = 0; ks < K / 16; ++ks) {
    const uint64_t dh = smem_desc(bhi + ks * 2 * N * 16, N * 16, 128);
    const uint32_t ah = tb + RC_A + 8 * ks;
    mma_ts(tb + RC_D, ah, dh, idesc, ks > 0);
    if (NPASS == 3) {
      const uint64_t dl = smem_desc(blo + ks * 2 * N * 16, N * 16, 128);
      mma_ts(tb + RC_D, tb + RC_ALO + 8 * ks, dh, idesc, true);
      mma_ts(tb + RC_D, ah, dl, idesc, true);
    }
  }
}
template <int NPASS, int LAYER>
__device__ __forceinline__ void r_issue_bwd(uint32_t tb, uint32_t sB_addr) {
  constexpr int NO = rl_n(LAYER), KI = rl_k(LAYER);
  constexpr uint32_t idesc = idesc_bf16(128, KI) | (1u << 16);
  const uint32_t bhi = sB_addr + rl_off(LAYER), blo = bhi + R_SET_BYTES;
#pragma unroll
  for (int ks = 0; ks < NO / 16; ++ks) {
    const uint64_t dh = smem_desc(bhi + ks * 256, 128, NO * 16);
    const uint32_t ah = tb + RC_A + 8 * ks;
    mma_ts(tb + RC_D, ah, dh, idesc, ks > 0);
    if (NPASS == 3) {
      const uint64_t dl = smem_desc(blo + ks * 256, 128, NO * 16);
      mma_ts(tb + RC_D, tb + RC_ALO + 8 * ks, dh, idesc, true);
      mma_ts(tb + RC_D, ah, dl, idesc, true);
    }
  }
}

#define NFB_RTC_SYNC_ISSUE(STMT)                                              \
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
#define NFB_RTC_FWD(LAYER) NFB_RTC_SYNC_ISSUE((r_issue_fwd<NPASS, LAYER>(tb, sB_addr)))
#define NFB_RTC_BWD(LAYER) NFB_RTC_SYNC_ISSUE((r_issue_bwd<NPASS, LAYER>(tb, sB_addr)))
#define NFB_RTC_WAIT()         \
  do {                         \
    mbar_wait(mbar, phase);    \
    phase ^= 1u;               \
    fence_after_sync();        \
  } while (0)

// log2-domain scores of one key row against a query: k4[d] = dimension d of the four heads; (head 0, head 1) in s01
__device__ __forceinline__ void attn_scores(const float4* __restrict__ k4, const float2 (&q01)[4], const float2 (&q23)[4],
                                            float2& s01, float2& s23) {
  const float4 a = k4[0], b = k4[1], c = k4[2], d = k4[3];
  s01 = __fmul2_rn(q01[0], make_float2(a.x, a.y));
  s23 = __fmul2_rn(q23[0], make_float2(a.z, a.w));
  s01 = __ffma2_rn(q01[1], make_float2(b.x, b.y), s01);
  s23 = __ffma2_rn(q23[1], make_float2(b.z, b.w), s23);
  s01 = __ffma2_rn(q01[2], make_float2(c.x, c.y), s01);
  s23 = __ffma2_rn(q23[2], make_float2(c.z, c.w), s23);
  s01 = __ffma2_rn(q01[3], make_float2(d.x, d.y), s01);
  s23 = __ffma2_rn(q23[3], make_float2(d.z, d.w), s23);
}

struct RayArgs {
  int R, S;
  const float* ps;
  const float* params;
  const float* pos_enc;
  float* raw;           // forward output
  const float* d_raw;   // backward input
  float* d_ps;          // backward output
  float* stash;         // forward (SAVE): out; stash backward: in
  uint8_t* pixel_mask;  // forward, optional: [R][S] = (number of valid observations > 1), render_ray.py:210
};

// ---- activation stash of the ray stage (forward SAVE -> k_ray_tc_bwd_stash): per 128-sample tile RP_PLANES planes of
// [128 samples][4 words], 560 B per sample ----
//   0..3 q (scaled, log2 domain, dimension-major, 0 for masked rows)   4..7 k   8..11 v (dimension-major)
//   12..15 o (normalised attention output)   16 -max   17 1/sum   18..21 xhat (LayerNorm)   22 rstd, z2, n_valid, -
//   23..30 ELU' codes of geometry_fc.0 (64)   31..32 of geometry_fc.2 (16)   33..34 of out_geometry_fc.0 (16)
enum : int { RP_Q = 0, RP_K = 4, RP_V = 8, RP_O = 12, RP_NM = 16, RP_IL = 17, RP_XHAT = 18, RP_MISC = 22, RP_H64 = 23,
             RP_G16 = 31, RP_HH = 33, RP_PLANES = 35 };
constexpr size_t RP_TILE_BYTES = (size_t)RP_PLANES * GROUP * 16;   // 71680

__device__ __forceinline__ void rp_st(float4* sp, int plane, float a, float b, float c, float d) {
  sp[plane * GROUP] = make_float4(a, b, c, d);
}
__device__ __forceinline__ void rp_st_codes8(float4* sp, int plane, const uint32_t (&q)[8]) {
  sp[plane * GROUP] = make_float4(__uint_as_float(q[0]), __uint_as_float(q[1]), __uint_as_float(q[2]), __uint_as_float(q[3]));
  sp[(plane + 1) * GROUP] = make_float4(__uint_as_float(q[4]), __uint_as_float(q[5]), __uint_as_float(q[6]), __uint_as_float(q[7]));
}
__device__ __forceinline__ void rp_ld_codes8(const float4* sp, int plane, uint32_t (&q)[8]) {
  const float4 a = __ldcs(sp + plane * GROUP), b = __ldcs(sp + (plane + 1) * GROUP);
  q[0] = __float_as_uint(a.x); q[1] = __float_as_uint(a.y); q[2] = __float_as_uint(a.z); q[3] = __float_as_uint(a.w);
  q[4] = __float_as_uint(b.x); q[5] = __float_as_uint(b.y); q[6] = __float_as_uint(b.z); q[7] = __float_as_uint(b.w);
}

template <int NPASS, bool BWD, bool SAVE>
__global__ void __launch_bounds__(GROUP * Cfg<NPASS, BWD>::NG, 1) k_ray_tc(RayArgs a) {
  using C = Cfg<NPASS, BWD>;
  constexpr int NG = C::NG;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sB = smem_raw;
  float* sf = reinterpret_cast<float*>(smem_raw + (size_t)R_SET_BYTES * (NPASS == 3 ? 2 : 1));
  float* s_pos = sf + RF_TOTAL;
  float* s_grp = s_pos + (size_t)a.S * 16;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_grp + (size_t)NG * C::GROUP_FLOATS);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + NG);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int grp = tid / GROUP, tg = tid % GROUP;
  // Rays longer than 128 samples are owned by a PAIR of groups (256 rows): the pair shares one K / V (/ Q / dO /
  // statistics) buffer of 256 rows and synchronises on its own 256-thread named barrier around the attention; the
  // dense layers stay per group (one M = 128 tile each).
  const bool pair_mode = a.S > GROUP;
  const int pair = grp >> 1, gp = grp & 1;
  const int AR = pair_mode ? 2 * GROUP : GROUP;                       // rows of the attention buffers
  float* abase = s_grp + (size_t)(pair_mode ? 2 * pair : grp) * C::GROUP_FLOATS;
  float* sk = abase;                                                  // [AR][16]
  float* sv = sk + AR * 16;
  float* sq = sv + AR * 16;                                           // [AR][16]
  float* sdo = sq + AR * 16;                                          // backward only
  float* sst = sdo + AR * 16;                                         // [AR][16]: -m | 1/l | -D | valid
  float* sx = sq + AR * 16;                                           // forward only: [AR][24] partial (l, o) of the partner's query
  float* skm = sx + AR * 24;                                          // forward only: [AR / 32][4] per-warp max |k_h|^2
  const int row = pair_mode ? gp * GROUP + tg : tg;                   // this thread's row in those buffers
  const int bar_id = 1 + grp;
  const int att_bar = pair_mode ? 9 + pair : bar_id, att_n = pair_mode ? 2 * GROUP : GROUP;
  uint64_t* mbar = s_bar + grp;

  if (warp == 0) tmem_alloc(s_tmem, NG * C::GC);
  if (tid == 0) {
    for (int g = 0; g < NG; ++g) mbar_init(s_bar + g, 1);
    mbar_init_fence();
  }
  {
    const float* p = a.params;
    load_rtile<NPASS>(sB, RL_GEO0, p + P_GEO0_W, 64, 65, tid, blockDim.x);
    load_rtile<NPASS>(sB, RL_GEO2, p + P_GEO2_W, 16, 64, tid, blockDim.x);
    load_rtile<NPASS>(sB, RL_QKV, p + P_ATT_Q, 48, 16, tid, blockDim.x);
    load_rtile<NPASS>(sB, RL_FC, p + P_ATT_FC, 16, 16, tid, blockDim.x);
    load_rtile<NPASS>(sB, RL_OG0, p + P_OG0_W, 16, 16, tid, blockDim.x);
    for (int i = tid; i < 64; i += blockDim.x) {
      sf[RF_B_GEO0 + i] = __ldg(p + P_GEO0_B + i);
      sf[RF_WCOL + i] = __ldg(p + P_GEO0_W + i * 65 + 64);
    }
    for (int i = tid; i < 16; i += blockDim.x) {
      sf[RF_B_GEO2 + i] = __ldg(p + P_GEO2_B + i);
      sf[RF_LNW + i] = __ldg(p + P_LN_W + i);
      sf[RF_LNB + i] = __ldg(p + P_LN_B + i);
      sf[RF_B_OG0 + i] = __ldg(p + P_OG0_B + i);
      sf[RF_W_OG2 + i] = __ldg(p + P_OG2_W + i);
    }
    if (tid == 0) sf[RF_B_OG2] = __ldg(p + P_OG2_B);
    for (int i = tid; i < a.S * 16; i += blockDim.x) s_pos[i] = __ldg(a.pos_enc + i);
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const uint32_t tb = *s_tmem + (uint32_t)(grp * C::GC);
  const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t sB_addr = smem_u32(sB);
  uint32_t phase = 0;

  const int S = a.S;
  const int RPG = pair_mode ? 1 : GROUP / S;       // whole rays per tile (>= 1)
  const int rl = pair_mode ? 0 : tg / S;           // ray within the tile
  const int s = pair_mode ? row : tg - rl * S;     // sample within the ray
  const int kb = (rl < RPG ? rl : 0) * S;          // first K / V row of this thread's ray
  const int ntiles = (a.R + RPG - 1) / RPG;
  const int tile0 = pair_mode ? blockIdx.x * (NG / 2) + pair : blockIdx.x * NG + grp;
  const int tstep = pair_mode ? gridDim.x * (NG / 2) : gridDim.x * NG;

  for (int tile = tile0; tile < ntiles; tile += tstep) {
    const int ray = tile * RPG + rl;
    const bool act = (rl < RPG) && (s < S) && (ray < a.R);
    const size_t smp = act ? ((size_t)ray * S + s) : 0;
    const float* psrow = a.ps + smp * NFB_PS_STRIDE;
    float4* sp = reinterpret_cast<float4*>(a.stash) + (size_t)(pair_mode ? 2 * tile + gp : tile) * (RP_PLANES * GROUP) + tg;
    const bool save = SAVE && act;

    // ---------------- geometry_fc.0 : 64 pooled statistics on the tensor cores, the mean weight on the CUDA cores ----
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      float t[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 q4 = __ldg(reinterpret_cast<const float4*>(psrow) + 4 * kc + j);
        t[4 * j] = q4.x; t[4 * j + 1] = q4.y; t[4 * j + 2] = q4.z; t[4 * j + 3] = q4.w;
      }
      r_put16<NPASS>(tl, kc, t);
    }
    const float4 tail = __ldg(reinterpret_cast<const float4*>(psrow) + 16);   // {mean weight, rgb}
    const float nvalid = __ldg(psrow + PS_NVALID);
    const bool row_valid = nvalid > 1.f;
    NFB_RTC_FWD(RL_GEO0);
    NFB_RTC_WAIT();
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      float h[16];
      r_ld16(tl, 16 * kc, h);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        h[j] = elu_fast(h[j] + fmaf(tail.x, sf[RF_WCOL + 16 * kc + j], sf[RF_B_GEO0 + 16 * kc + j]));
      if (BWD || SAVE) {
        uint32_t q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) q[j] = elu_stash_pack(h[2 * j], h[2 * j + 1]);
        if (BWD) tmem_st8(tl + RC_HQ + 8 * kc, q);
        if (save) rp_st_codes8(sp, RP_H64 + 2 * kc, q);
      }
      r_put16<NPASS>(tl, kc, h);
    }
    NFB_RTC_FWD(RL_GEO2);
    NFB_RTC_WAIT();
    float xin[16];
    uint32_t gq[8];
    {
      r_ld16(tl, 0, xin);
#pragma unroll
      for (int j = 0; j < 16; ++j) xin[j] = elu_fast(xin[j] + sf[RF_B_GEO2 + j]);
      if (BWD || SAVE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) gq[j] = elu_stash_pack(xin[2 * j], xin[2 * j + 1]);
        if (save) rp_st_codes8(sp, RP_G16, gq);
      }
      const float* pe = s_pos + (act ? s : 0) * 16;
#pragma unroll
      for (int j = 0; j < 16; ++j) xin[j] += pe[j];
      r_put16<NPASS>(tl, 0, xin);
    }
    NFB_RTC_FWD(RL_QKV);
    NFB_RTC_WAIT();

    // ---------------- attention ----------------
    // K / V / Q rows live in shared memory DIMENSION-major ([d][head]): one LDS.128 brings dimension d of all four
    // heads, and the (head 0, head 1) / (head 2, head 3) pairs feed the packed FFMA2 / FADD2 / FMUL2 pipes (on sm_100
    // the scalar fp32 instructions issue at half the packed rate).  Masked query rows (<= 1 valid view:
    // masked_fill(mask == 0, -1e9) on the whole row = uniform attention) carry q = 0, so their scores, maxima and
    // probabilities come out as 0, 0 and 1 without a branch.
    // TWO-QUERY form (forward kernel, S a multiple of 64): the broadcast K / V row loads are the bound of the loop above (12
    // LDS.128 per key and query over two passes: ncu shows the shared-memory pipe at 70 %).  Warps w and w ^ 1 of a ray split
    // its keys in halves; every thread runs its OWN query and the query of the same lane of the partner warp over its half,
    // so a K / V row is loaded once per TWO (query, key) pairs, and the partial (l, o) of the partner's query is handed over
    // through shared memory.  The max pass disappears: softmax is shift invariant, and |q_h| max_j |k_jh| >= max_j q_h.k_jh
    // (Cauchy-Schwarz) is a valid shift that needs no pass over the keys; 2^(s - shift) cannot underflow the whole row while
    // the shift is < 60 (log2 units) -- beyond that (never seen; scores of +-40 nats) the warp pair falls back to the exact
    // two-pass loop for its own queries.  4 LDS.128 per (query, key) instead of 12, 28 instructions instead of 48.
    const bool fast = !BWD && (S % 64 == 0);
    float2 q01[4], q23[4];
    float o[16], m2[4], il[4];
    if (fast) {
      float kn[4];
      {
        float qq[16], kk[16], vv[16];
        r_ld16(tl, 0, qq);
        r_ld16(tl, 16, kk);
        r_ld16(tl, 32, vv);
        const float qs = row_valid ? INV_TEMP * LOG2E : 0.f;          // scores in the log2 domain
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          q01[d] = make_float2(qq[d] * qs, qq[4 + d] * qs);
          q23[d] = make_float2(qq[8 + d] * qs, qq[12 + d] * qs);
          *reinterpret_cast<float4*>(sk + row * 16 + 4 * d) = make_float4(kk[d], kk[4 + d], kk[8 + d], kk[12 + d]);
          *reinterpret_cast<float4*>(sv + row * 16 + 4 * d) = make_float4(vv[d], vv[4 + d], vv[8 + d], vv[12 + d]);
          *reinterpret_cast<float4*>(sq + row * 16 + 4 * d) = make_float4(q01[d].x, q01[d].y, q23[d].x, q23[d].y);
          if (save) {
            rp_st(sp, RP_Q + d, q01[d].x, q01[d].y, q23[d].x, q23[d].y);
            rp_st(sp, RP_K + d, kk[d], kk[4 + d], kk[8 + d], kk[12 + d]);
            rp_st(sp, RP_V + d, vv[d], vv[4 + d], vv[8 + d], vv[12 + d]);
          }
        }
#pragma unroll
        for (int h = 0; h < 4; ++h)
          kn[h] = fmaf(kk[4 * h + 3], kk[4 * h + 3], fmaf(kk[4 * h + 2], kk[4 * h + 2], fmaf(kk[4 * h + 1], kk[4 * h + 1], kk[4 * h] * kk[4 * h])));
      }
      {                                               // residual input parked in the free D columns while the loop runs
        uint32_t park[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) park[c] = __float_as_uint(xin[c]);
        tmem_st16(tl + RC_D + 48, park);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int h = 0; h < 4; ++h) kn[h] = fmaxf(kn[h], __shfl_xor_sync(0xffffffffu, kn[h], off));
      }
      if ((tid & 31) == 0) *reinterpret_cast<float4*>(skm + (row >> 5) * 4) = make_float4(kn[0], kn[1], kn[2], kn[3]);
      named_bar_sync(att_bar, att_n);
      float4 km = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int wv = kb >> 5; wv < ((kb + S) >> 5); ++wv) {
        const float4 t4 = *reinterpret_cast<const float4*>(skm + wv * 4);
        km.x = fmaxf(km.x, t4.x); km.y = fmaxf(km.y, t4.y); km.z = fmaxf(km.z, t4.z); km.w = fmaxf(km.w, t4.w);
      }
      const int prow = row ^ 32;                       // same lane of the partner warp (same ray: S is a multiple of 64)
      float2 qb01[4], qb23[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const float4 t4 = *reinterpret_cast<const float4*>(sq + prow * 16 + 4 * d);
        qb01[d] = make_float2(t4.x, t4.y);
        qb23[d] = make_float2(t4.z, t4.w);
      }
      // shift of a query: sqrt(|q_h|^2 max_j |k_jh|^2), inflated by 1e-4 relative + 1e-6 so that rounding cannot put a score above it.
      // Both threads of a pair evaluate this expression on the same numbers in the same order: bit-identical shifts.
      float shA[4], shB[4];
      {
        const float kmx[4] = {km.x, km.y, km.z, km.w};
        float qa[4] = {0.f, 0.f, 0.f, 0.f}, qb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          qa[0] = fmaf(q01[d].x, q01[d].x, qa[0]); qa[1] = fmaf(q01[d].y, q01[d].y, qa[1]);
          qa[2] = fmaf(q23[d].x, q23[d].x, qa[2]); qa[3] = fmaf(q23[d].y, q23[d].y, qa[3]);
          qb[0] = fmaf(qb01[d].x, qb01[d].x, qb[0]); qb[1] = fmaf(qb01[d].y, qb01[d].y, qb[1]);
          qb[2] = fmaf(qb23[d].x, qb23[d].x, qb[2]); qb[3] = fmaf(qb23[d].y, qb23[d].y, qb[3]);
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          shA[h] = fmaf(sqrtf(qa[h] * kmx[h]), 1.0001f, 1e-6f);
          shB[h] = fmaf(sqrtf(qb[h] * kmx[h]), 1.0001f, 1e-6f);
        }
      }
      bool big = false;
#pragma unroll
      for (int h = 0; h < 4; ++h) big = big || !(shA[h] < 60.f) || !(shB[h] < 60.f);
      float2 lB01 = make_float2(0.f, 0.f), lB23 = lB01;
      float2 oB01[4], oB23[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) oB01[d] = oB23[d] = make_float2(0.f, 0.f);
      float2 l01 = make_float2(0.f, 0.f), l23 = l01;
      float2 o01[4], o23[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) o01[d] = o23[d] = make_float2(0.f, 0.f);
      if (__any_sync(0xffffffffu, big)) {
        // exact form for this warp pair (the partner warp takes the same branch: it sees the same two shift sets): own query
        // over ALL keys with the true row maximum; the partner's partial stays zero
        const float4* kr = reinterpret_cast<const float4*>(sk + kb * 16);
        const float4* vr = reinterpret_cast<const float4*>(sv + kb * 16);
        float2 mx01 = make_float2(-3.4e38f, -3.4e38f), mx23 = mx01;
        for (int j = 0; j < S; ++j) {
          float2 s01, s23;
          attn_scores(kr + 4 * j, q01, q23, s01, s23);
          mx01.x = fmaxf(mx01.x, s01.x); mx01.y = fmaxf(mx01.y, s01.y);
          mx23.x = fmaxf(mx23.x, s23.x); mx23.y = fmaxf(mx23.y, s23.y);
        }
        shA[0] = mx01.x; shA[1] = mx01.y; shA[2] = mx23.x; shA[3] = mx23.y;
        const float2 nm01 = make_float2(-mx01.x, -mx01.y), nm23 = make_float2(-mx23.x, -mx23.y);
        for (int j = 0; j < S; ++j) {
          float2 s01, s23;
          attn_scores(kr + 4 * j, q01, q23, s01, s23);
          s01 = __fadd2_rn(s01, nm01);
          s23 = __fadd2_rn(s23, nm23);
          const float2 p01 = make_float2(ex2_approx(s01.x), ex2_approx(s01.y));
          const float2 p23 = make_float2(ex2_approx(s23.x), ex2_approx(s23.y));
          l01 = __fadd2_rn(l01, p01);
          l23 = __fadd2_rn(l23, p23);
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            const float4 vj = vr[4 * j + d];
            o01[d] = __ffma2_rn(p01, make_float2(vj.x, vj.y), o01[d]);
            o23[d] = __ffma2_rn(p23, make_float2(vj.z, vj.w), o23[d]);
          }
        }
      } else {
        const int j0 = kb + ((row >> 5) & 1) * (S >> 1);
        const float4* kr = reinterpret_cast<const float4*>(sk + j0 * 16);
        const float4* vr = reinterpret_cast<const float4*>(sv + j0 * 16);
        const float2 nA01 = make_float2(-shA[0], -shA[1]), nA23 = make_float2(-shA[2], -shA[3]);
        const float2 nB01 = make_float2(-shB[0], -shB[1]), nB23 = make_float2(-shB[2], -shB[3]);
        const int half = S >> 1;
#pragma unroll 2
        for (int j = 0; j < half; ++j) {
          float2 sa01, sa23, sb01, sb23;
          {
            const float4 ka = kr[4 * j], kb4 = kr[4 * j + 1], kc = kr[4 * j + 2], kd = kr[4 * j + 3];
            const float2 a01 = make_float2(ka.x, ka.y), a23 = make_float2(ka.z, ka.w);
            const float2 b01 = make_float2(kb4.x, kb4.y), b23 = make_float2(kb4.z, kb4.w);
            const float2 c01 = make_float2(kc.x, kc.y), c23 = make_float2(kc.z, kc.w);
            const float2 d01 = make_float2(kd.x, kd.y), d23 = make_float2(kd.z, kd.w);
            sa01 = __ffma2_rn(q01[0], a01, nA01); sa23 = __ffma2_rn(q23[0], a23, nA23);
            sb01 = __ffma2_rn(qb01[0], a01, nB01); sb23 = __ffma2_rn(qb23[0], a23, nB23);
            sa01 = __ffma2_rn(q01[1], b01, sa01); sa23 = __ffma2_rn(q23[1], b23, sa23);
            sb01 = __ffma2_rn(qb01[1], b01, sb01); sb23 = __ffma2_rn(qb23[1], b23, sb23);
            sa01 = __ffma2_rn(q01[2], c01, sa01); sa23 = __ffma2_rn(q23[2], c23, sa23);
            sb01 = __ffma2_rn(qb01[2], c01, sb01); sb23 = __ffma2_rn(qb23[2], c23, sb23);
            sa01 = __ffma2_rn(q01[3], d01, sa01); sa23 = __ffma2_rn(q23[3], d23, sa23);
            sb01 = __ffma2_rn(qb01[3], d01, sb01); sb23 = __ffma2_rn(qb23[3], d23, sb23);
          }
          const float2 pa01 = make_float2(ex2_approx(sa01.x), ex2_approx(sa01.y));
          const float2 pa23 = make_float2(ex2_approx(sa23.x), ex2_approx(sa23.y));
          const float2 pb01 = make_float2(ex2_approx(sb01.x), ex2_approx(sb01.y));
          const float2 pb23 = make_float2(ex2_approx(sb23.x), ex2_approx(sb23.y));
          l01 = __fadd2_rn(l01, pa01); l23 = __fadd2_rn(l23, pa23);
          lB01 = __fadd2_rn(lB01, pb01); lB23 = __fadd2_rn(lB23, pb23);
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            const float4 vj = vr[4 * j + d];
            const float2 v01 = make_float2(vj.x, vj.y), v23 = make_float2(vj.z, vj.w);
            o01[d] = __ffma2_rn(pa01, v01, o01[d]);
            o23[d] = __ffma2_rn(pa23, v23, o23[d]);
            oB01[d] = __ffma2_rn(pb01, v01, oB01[d]);
            oB23[d] = __ffma2_rn(pb23, v23, oB23[d]);
          }
        }
      }
      // hand the partner's partial over, collect mine
      {
        float4* xo = reinterpret_cast<float4*>(sx + prow * 24);
        xo[0] = make_float4(lB01.x, lB01.y, lB23.x, lB23.y);
#pragma unroll
        for (int d = 0; d < 4; ++d) xo[1 + d] = make_float4(oB01[d].x, oB01[d].y, oB23[d].x, oB23[d].y);
      }
      named_bar_sync(att_bar, att_n);
      {
        const float4* xi = reinterpret_cast<const float4*>(sx + row * 24);
        const float4 lp = xi[0];
        l01 = __fadd2_rn(l01, make_float2(lp.x, lp.y));
        l23 = __fadd2_rn(l23, make_float2(lp.z, lp.w));
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const float4 op = xi[1 + d];
          o01[d] = __fadd2_rn(o01[d], make_float2(op.x, op.y));
          o23[d] = __fadd2_rn(o23[d], make_float2(op.z, op.w));
        }
      }
      il[0] = 1.f / l01.x; il[1] = 1.f / l01.y; il[2] = 1.f / l23.x; il[3] = 1.f / l23.y;
#pragma unroll
      for (int h = 0; h < 4; ++h) m2[h] = shA[h];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        o[d] = o01[d].x * il[0];
        o[4 + d] = o01[d].y * il[1];
        o[8 + d] = o23[d].x * il[2];
        o[12 + d] = o23[d].y * il[3];
      }
      {
        uint32_t park[16];
        tmem_ld16u(tl + RC_D + 48, park);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) xin[c] = __uint_as_float(park[c]);
      }
    } else {
    {
      float qq[16], kk[16], vv[16];
      r_ld16(tl, 0, qq);
      r_ld16(tl, 16, kk);
      r_ld16(tl, 32, vv);
      const float qs = row_valid ? INV_TEMP * LOG2E : 0.f;          // scores in the log2 domain
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        q01[d] = make_float2(qq[d] * qs, qq[4 + d] * qs);
        q23[d] = make_float2(qq[8 + d] * qs, qq[12 + d] * qs);
        *reinterpret_cast<float4*>(sk + row * 16 + 4 * d) = make_float4(kk[d], kk[4 + d], kk[8 + d], kk[12 + d]);
        *reinterpret_cast<float4*>(sv + row * 16 + 4 * d) = make_float4(vv[d], vv[4 + d], vv[8 + d], vv[12 + d]);
        if (save) {
          rp_st(sp, RP_Q + d, q01[d].x, q01[d].y, q23[d].x, q23[d].y);
          rp_st(sp, RP_K + d, kk[d], kk[4 + d], kk[8 + d], kk[12 + d]);
          rp_st(sp, RP_V + d, vv[d], vv[4 + d], vv[8 + d], vv[12 + d]);
        }
      }
    }
    named_bar_sync(att_bar, att_n);
    {
      const float4* kr = reinterpret_cast<const float4*>(sk + kb * 16);
      const float4* vr = reinterpret_cast<const float4*>(sv + kb * 16);
      float2 mx01 = make_float2(-3.4e38f, -3.4e38f), mx23 = mx01;
      for (int j = 0; j < S; ++j) {
        float2 s01, s23;
        attn_scores(kr + 4 * j, q01, q23, s01, s23);
        mx01.x = fmaxf(mx01.x, s01.x); mx01.y = fmaxf(mx01.y, s01.y);
        mx23.x = fmaxf(mx23.x, s23.x); mx23.y = fmaxf(mx23.y, s23.y);
      }
      m2[0] = mx01.x; m2[1] = mx01.y; m2[2] = mx23.x; m2[3] = mx23.y;
      const float2 nm01 = make_float2(-mx01.x, -mx01.y), nm23 = make_float2(-mx23.x, -mx23.y);
      float2 l01 = make_float2(0.f, 0.f), l23 = l01;
      float2 o01[4], o23[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) o01[d] = o23[d] = make_float2(0.f, 0.f);
      for (int j = 0; j < S; ++j) {
        float2 s01, s23;
        attn_scores(kr + 4 * j, q01, q23, s01, s23);
        s01 = __fadd2_rn(s01, nm01);
        s23 = __fadd2_rn(s23, nm23);
        const float2 p01 = make_float2(ex2_approx(s01.x), ex2_approx(s01.y));
        const float2 p23 = make_float2(ex2_approx(s23.x), ex2_approx(s23.y));
        l01 = __fadd2_rn(l01, p01);
        l23 = __fadd2_rn(l23, p23);
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const float4 vj = vr[4 * j + d];
          o01[d] = __ffma2_rn(p01, make_float2(vj.x, vj.y), o01[d]);
          o23[d] = __ffma2_rn(p23, make_float2(vj.z, vj.w), o23[d]);
        }
      }
      il[0] = 1.f / l01.x; il[1] = 1.f / l01.y; il[2] = 1.f / l23.x; il[3] = 1.f / l23.y;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        o[d] = o01[d].x * il[0];
        o[4 + d] = o01[d].y * il[1];
        o[8 + d] = o23[d].x * il[2];
        o[12 + d] = o23[d].y * il[3];
      }
    }
    }

    if (save) {
#pragma unroll
      for (int j = 0; j < 4; ++j) rp_st(sp, RP_O + j, o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      rp_st(sp, RP_NM, -m2[0], -m2[1], -m2[2], -m2[3]);
      rp_st(sp, RP_IL, il[0], il[1], il[2], il[3]);
    }
    // ---------------- fc + residual + LayerNorm + sigma head ----------------
    r_put16<NPASS>(tl, 0, o);
    NFB_RTC_FWD(RL_FC);
    NFB_RTC_WAIT();
    float xhat[16];
    float rstd;
    {
      float y[16];
      r_ld16(tl, 0, y);
      float mu = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        y[c] += xin[c];
        mu += y[c];
      }
      mu *= (1.f / 16.f);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) var = fmaf(y[c] - mu, y[c] - mu, var);
      var *= (1.f / 16.f);
      rstd = rsqrtf(var + LN_EPS);
      float ln[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        xhat[c] = (y[c] - mu) * rstd;
        ln[c] = fmaf(xhat[c], sf[RF_LNW + c], sf[RF_LNB + c]);
      }
      if (save) {
#pragma unroll
        for (int j = 0; j < 4; ++j) rp_st(sp, RP_XHAT + j, xhat[4 * j], xhat[4 * j + 1], xhat[4 * j + 2], xhat[4 * j + 3]);
      }
      r_put16<NPASS>(tl, 0, ln);
    }
    NFB_RTC_FWD(RL_OG0);
    NFB_RTC_WAIT();
    float hh[16];
    float z2 = sf[RF_B_OG2];
    r_ld16(tl, 0, hh);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      hh[c] = elu_fast(hh[c] + sf[RF_B_OG0 + c]);
      z2 = fmaf(hh[c], sf[RF_W_OG2 + c], z2);
    }

    if (save) {
      uint32_t hq[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) hq[j] = elu_stash_pack(hh[2 * j], hh[2 * j + 1]);
      rp_st_codes8(sp, RP_HH, hq);
      rp_st(sp, RP_MISC, rstd, z2, nvalid, 0.f);
    }
    if (!BWD) {
      float sigma = fmaxf(z2, 0.f);
      if (nvalid < 1.f) sigma = 0.f;                     // mlp_network.py:265
      if (act) {
        reinterpret_cast<float4*>(a.raw)[smp] = make_float4(tail.y, tail.z, tail.w, sigma);
        if (a.pixel_mask) a.pixel_mask[smp] = nvalid > 1.f ? 1 : 0;
      }
      named_bar_sync(att_bar, att_n);                    // K / V rows are rewritten by the next tile
      continue;
    }

    // =================================== backward ===================================
    if (BWD) {
      float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
      if (act) dr = __ldg(reinterpret_cast<const float4*>(a.d_raw) + smp);
      const float dz2 = (z2 > 0.f && !(nvalid < 1.f)) ? dr.w : 0.f;
      // sigma head
      {
        float dh[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) dh[k] = dz2 * sf[RF_W_OG2 + k] * elu_grad_from_out(hh[k]);
        r_put16<NPASS>(tl, 0, dh);
      }
      NFB_RTC_BWD(RL_OG0);
      NFB_RTC_WAIT();
      // LayerNorm backward: dy = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dln * gamma
      float dy[16];
      {
        r_ld16(tl, 0, dy);
        float gsum = 0.f, gx = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          dy[c] *= sf[RF_LNW + c];
          gsum += dy[c];
          gx = fmaf(dy[c], xhat[c], gx);
        }
        gsum *= (1.f / 16.f); gx *= (1.f / 16.f);
#pragma unroll
        for (int c = 0; c < 16; ++c) dy[c] = rstd * (dy[c] - gsum - xhat[c] * gx);
        r_put16<NPASS>(tl, 0, dy);
      }
      NFB_RTC_BWD(RL_FC);
      NFB_RTC_WAIT();
      float2 dO01[4], dO23[4];
      float2 Dh01, Dh23;
      {
        float dO[16];
        r_ld16(tl, 0, dO);
        float Dh[4];
#pragma unroll
        for (int h = 0; h < 4; ++h)
          Dh[h] = dO[4 * h] * o[4 * h] + dO[4 * h + 1] * o[4 * h + 1] + dO[4 * h + 2] * o[4 * h + 2] + dO[4 * h + 3] * o[4 * h + 3];
        Dh01 = make_float2(Dh[0], Dh[1]);
        Dh23 = make_float2(Dh[2], Dh[3]);
        // publish per-query quantities for the key-side pass (dimension-major like K / V)
        const float vf = row_valid ? 1.f : 0.f;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          dO01[d] = make_float2(dO[d], dO[4 + d]);
          dO23[d] = make_float2(dO[8 + d], dO[12 + d]);
          *reinterpret_cast<float4*>(sq + row * 16 + 4 * d) = make_float4(q01[d].x, q01[d].y, q23[d].x, q23[d].y);
          *reinterpret_cast<float4*>(sdo + row * 16 + 4 * d) = make_float4(dO[d], dO[4 + d], dO[8 + d], dO[12 + d]);
        }
        *reinterpret_cast<float4*>(sst + row * 16) = make_float4(-m2[0], -m2[1], -m2[2], -m2[3]);
        *reinterpret_cast<float4*>(sst + row * 16 + 4) = make_float4(il[0], il[1], il[2], il[3]);
        *reinterpret_cast<float4*>(sst + row * 16 + 8) = make_float4(-Dh[0], -Dh[1], -Dh[2], -Dh[3]);
        *reinterpret_cast<float4*>(sst + row * 16 + 12) = make_float4(vf, vf, vf, vf);
      }
      named_bar_sync(att_bar, att_n);

      float dqkv[48];
      // query side: dq_i = sum_j dS_ij k_j  (zero for masked rows: masked_fill blocks the gradient)
      {
        const float4* kr = reinterpret_cast<const float4*>(sk + kb * 16);
        const float4* vr = reinterpret_cast<const float4*>(sv + kb * 16);
        const float2 nm01 = make_float2(-m2[0], -m2[1]), nm23 = make_float2(-m2[2], -m2[3]);
        const float2 il01 = make_float2(il[0], il[1]), il23 = make_float2(il[2], il[3]);
        const float2 nD01 = make_float2(-Dh01.x, -Dh01.y), nD23 = make_float2(-Dh23.x, -Dh23.y);
        float2 dq01[4], dq23[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) dq01[d] = dq23[d] = make_float2(0.f, 0.f);
        for (int j = 0; j < S; ++j) {
          float2 s01, s23;
          attn_scores(kr + 4 * j, q01, q23, s01, s23);
          s01 = __fadd2_rn(s01, nm01);
          s23 = __fadd2_rn(s23, nm23);
          const float2 p01 = __fmul2_rn(make_float2(ex2_approx(s01.x), ex2_approx(s01.y)), il01);
          const float2 p23 = __fmul2_rn(make_float2(ex2_approx(s23.x), ex2_approx(s23.y)), il23);
          float2 dP01 = nD01, dP23 = nD23;                  // dP - D
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            const float4 vj = vr[4 * j + d];
            dP01 = __ffma2_rn(dO01[d], make_float2(vj.x, vj.y), dP01);
            dP23 = __ffma2_rn(dO23[d], make_float2(vj.z, vj.w), dP23);
          }
          const float2 dS01 = __fmul2_rn(p01, dP01), dS23 = __fmul2_rn(p23, dP23);
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            const float4 kj = kr[4 * j + d];
            dq01[d] = __ffma2_rn(dS01, make_float2(kj.x, kj.y), dq01[d]);
            dq23[d] = __ffma2_rn(dS23, make_float2(kj.z, kj.w), dq23[d]);
          }
        }
        const float sc = row_valid ? INV_TEMP : 0.f;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          dqkv[d] = dq01[d].x * sc;
          dqkv[4 + d] = dq01[d].y * sc;
          dqkv[8 + d] = dq23[d].x * sc;
          dqkv[12 + d] = dq23[d].y * sc;
        }
      }
      // key side: dk_j = sum_i dS_ij q_i ; dv_j = sum_i p_ij dO_i   (this thread is key j)
      {
        float2 k01[4], k23[4], v01[4], v23[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const float4 k4 = *reinterpret_cast<const float4*>(sk + row * 16 + 4 * d);
          const float4 v4 = *reinterpret_cast<const float4*>(sv + row * 16 + 4 * d);
          k01[d] = make_float2(k4.x, k4.y); k23[d] = make_float2(k4.z, k4.w);
          v01[d] = make_float2(v4.x, v4.y); v23[d] = make_float2(v4.z, v4.w);
        }
        float2 dk01[4], dk23[4], dv01[4], dv23[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) dk01[d] = dk23[d] = dv01[d] = dv23[d] = make_float2(0.f, 0.f);
        const float4* qr = reinterpret_cast<const float4*>(sq + kb * 16);
        const float4* gr = reinterpret_cast<const float4*>(sdo + kb * 16);
        const float4* st = reinterpret_cast<const float4*>(sst + kb * 16);
        for (int i = 0; i < S; ++i) {
          float4 qi[4], gi[4];
#pragma unroll
          for (int d = 0; d < 4; ++d) { qi[d] = qr[4 * i + d]; gi[d] = gr[4 * i + d]; }
          const float4 nm = st[4 * i], ili = st[4 * i + 1], nD = st[4 * i + 2], vf = st[4 * i + 3];
          // masked query rows carry q = 0, -m = 0 and 1/l = 1/S: p = 1/S as the reference's uniform row, dS forced to 0
          float2 s01 = make_float2(nm.x, nm.y), s23 = make_float2(nm.z, nm.w);
          float2 dP01 = make_float2(nD.x, nD.y), dP23 = make_float2(nD.z, nD.w);
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            s01 = __ffma2_rn(make_float2(qi[d].x, qi[d].y), k01[d], s01);
            s23 = __ffma2_rn(make_float2(qi[d].z, qi[d].w), k23[d], s23);
            dP01 = __ffma2_rn(make_float2(gi[d].x, gi[d].y), v01[d], dP01);
            dP23 = __ffma2_rn(make_float2(gi[d].z, gi[d].w), v23[d], dP23);
          }
          const float2 p01 = __fmul2_rn(make_float2(ex2_approx(s01.x), ex2_approx(s01.y)), make_float2(ili.x, ili.y));
          const float2 p23 = __fmul2_rn(make_float2(ex2_approx(s23.x), ex2_approx(s23.y)), make_float2(ili.z, ili.w));
          const float2 dS01 = __fmul2_rn(__fmul2_rn(p01, dP01), make_float2(vf.x, vf.y));
          const float2 dS23 = __fmul2_rn(__fmul2_rn(p23, dP23), make_float2(vf.z, vf.w));
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            dk01[d] = __ffma2_rn(dS01, make_float2(qi[d].x, qi[d].y), dk01[d]);
            dk23[d] = __ffma2_rn(dS23, make_float2(qi[d].z, qi[d].w), dk23[d]);
            dv01[d] = __ffma2_rn(p01, make_float2(gi[d].x, gi[d].y), dv01[d]);
            dv23[d] = __ffma2_rn(p23, make_float2(gi[d].z, gi[d].w), dv23[d]);
          }
        }
        // q in shared memory carries the 0.5 * log2(e) score scale; d k = sum dS * q_scaled / log2(e)
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          dqkv[16 + d] = dk01[d].x * (1.f / LOG2E);
          dqkv[16 + 4 + d] = dk01[d].y * (1.f / LOG2E);
          dqkv[16 + 8 + d] = dk23[d].x * (1.f / LOG2E);
          dqkv[16 + 12 + d] = dk23[d].y * (1.f / LOG2E);
          dqkv[32 + d] = dv01[d].x;
          dqkv[32 + 4 + d] = dv01[d].y;
          dqkv[32 + 8 + d] = dv23[d].x;
          dqkv[32 + 12 + d] = dv23[d].y;
        }
      }
      // d xin = dy (residual) + [dq | dk | dv] Wqkv ; pos_encoding is a constant
#pragma unroll
      for (int kc = 0; kc < 3; ++kc) {
        float t[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) t[j] = dqkv[16 * kc + j];
        r_put16<NPASS>(tl, kc, t);
      }
      NFB_RTC_BWD(RL_QKV);
      NFB_RTC_WAIT();
      {
        float dx[16];
        r_ld16(tl, 0, dx);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dx[2 * j] = (dx[2 * j] + dy[2 * j]) * elu_stash_lo(gq[j]);
          dx[2 * j + 1] = (dx[2 * j + 1] + dy[2 * j + 1]) * elu_stash_hi(gq[j]);
        }
        r_put16<NPASS>(tl, 0, dx);
      }
      NFB_RTC_BWD(RL_GEO2);
      NFB_RTC_WAIT();
      float dwm = 0.f;
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        float dh[16];
        r_ld16(tl, 16 * kc, dh);
        uint32_t qc[8];
        tmem_ld8u(tl + RC_HQ + 8 * kc, qc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dh[2 * j] *= elu_stash_lo(qc[j]);
          dh[2 * j + 1] *= elu_stash_hi(qc[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) dwm = fmaf(dh[j], sf[RF_WCOL + 16 * kc + j], dwm);
        r_put16<NPASS>(tl, kc, dh);
      }
      NFB_RTC_BWD(RL_GEO0);
      NFB_RTC_WAIT();
      {
        float4* out = reinterpret_cast<float4*>(a.d_ps + smp * NFB_PS_STRIDE);
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          float t[16];
          r_ld16(tl, 16 * kc, t);
          if (act) {
#pragma unroll
            for (int j = 0; j < 4; ++j) out[4 * kc + j] = make_float4(t[4 * j], t[4 * j + 1], t[4 * j + 2], t[4 * j + 3]);
          }
        }
        if (act) {
          out[16] = make_float4(dwm, dr.x, dr.y, dr.z);
          out[17] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      named_bar_sync(att_bar, att_n);                    // K / V / Q / dO rows are rewritten by the next tile
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*s_tmem, NG * C::GC);
}

template <int NPASS, bool BWD, bool SAVE = false>
int launch_ray_tc(const RayArgs& a, cudaStream_t st) {
  using C = Cfg<NPASS, BWD>;
  const size_t smem = C::smem(a.S);
  cudaError_t e = cudaFuncSetAttribute(k_ray_tc<NPASS, BWD, SAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "k_ray_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const bool pair_mode = a.S > GROUP;
  const int RPG = pair_mode ? 1 : GROUP / a.S;
  const int ntiles = (a.R + RPG - 1) / RPG;
  const int per_cta = pair_mode ? C::NG / 2 : C::NG;
  int grid = (ntiles + per_cta - 1) / per_cta;
  const int cap = nfb_num_sms();
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  k_ray_tc<NPASS, BWD, SAVE><<<grid, GROUP * C::NG, smem, st>>>(a);
  NFB_CHECK_LAUNCH("k_ray_tc");
  return NFB_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Data-gradient of the ray stage FROM THE STASH: no forward recompute, 128 registers, 4 groups (16 warps) per SM.
// TMEM columns of a group (128): D [0,64) | A hi [64,96) | A lo [96,128); dy is parked in D columns [48,64) while the
// attention passes run.
// ---------------------------------------------------------------------------------------------------------------
constexpr int BS_NG = 4;
constexpr int BS_GROUP_FLOATS = 4 * GROUP * 16 + GROUP * 16;    // K, V, Q, dO rows + statistics
template <int NPASS>
size_t bwd_stash_smem() {
  return (size_t)R_SET_BYTES * (NPASS == 3 ? 2 : 1) + sizeof(float) * (RF_TOTAL + (size_t)BS_NG * BS_GROUP_FLOATS) + BS_NG * 8 + 16;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int NPASS>
__global__ void __launch_bounds__(GROUP * BS_NG, 1) k_ray_tc_bwd_stash(RayArgs a) {
  constexpr int NG = BS_NG;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sB = smem_raw;
  float* sf = reinterpret_cast<float*>(smem_raw + (size_t)R_SET_BYTES * (NPASS == 3 ? 2 : 1));
  float* s_grp = sf + RF_TOTAL;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_grp + (size_t)NG * BS_GROUP_FLOATS);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + NG);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int grp = tid / GROUP, tg = tid % GROUP;
  const bool pair_mode = a.S > GROUP;                   // rays of 129..256 samples: two groups per ray (see k_ray_tc)
  const int pair = grp >> 1, gp = grp & 1;
  const int AR = pair_mode ? 2 * GROUP : GROUP;
  float* abase = s_grp + (size_t)(pair_mode ? 2 * pair : grp) * BS_GROUP_FLOATS;
  float* sk = abase;                                    // [AR][16] each, dimension-major rows
  float* sv = sk + AR * 16;
  float* sq = sv + AR * 16;
  float* sdo = sq + AR * 16;
  float* sst = sdo + AR * 16;                           // [AR][16]: -m[4] | 1/l[4] | -D[4] | valid x4
  const int row = pair_mode ? gp * GROUP + tg : tg;
  const int bar_id = 1 + grp;
  const int att_bar = pair_mode ? 9 + pair : bar_id, att_n = pair_mode ? 2 * GROUP : GROUP;
  uint64_t* mbar = s_bar + grp;

  if (warp == 0) tmem_alloc(s_tmem, NG * 128);
  if (tid == 0) {
    for (int g = 0; g < NG; ++g) mbar_init(s_bar + g, 1);
    mbar_init_fence();
  }
  {
    const float* p = a.params;
    load_rtile<NPASS>(sB, RL_GEO0, p + P_GEO0_W, 64, 65, tid, blockDim.x);
    load_rtile<NPASS>(sB, RL_GEO2, p + P_GEO2_W, 16, 64, tid, blockDim.x);
    load_rtile<NPASS>(sB, RL_QKV, p + P_ATT_Q, 48, 16, tid, blockDim.x);
    load_rtile<NPASS>(sB, RL_FC, p + P_ATT_FC, 16, 16, tid, blockDim.x);
    load_rtile<NPASS>(sB, RL_OG0, p + P_OG0_W, 16, 16, tid, blockDim.x);
    for (int i = tid; i < 64; i += blockDim.x) {
      sf[RF_B_GEO0 + i] = __ldg(p + P_GEO0_B + i);
      sf[RF_WCOL + i] = __ldg(p + P_GEO0_W + i * 65 + 64);
    }
    for (int i = tid; i < 16; i += blockDim.x) {
      sf[RF_LNW + i] = __ldg(p + P_LN_W + i);
      sf[RF_W_OG2 + i] = __ldg(p + P_OG2_W + i);
    }
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const uint32_t tb = *s_tmem + (uint32_t)(grp * 128);
  const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t sB_addr = smem_u32(sB);
  uint32_t phase = 0;

  const int S = a.S;
  const int RPG = pair_mode ? 1 : GROUP / S;
  const int rl = pair_mode ? 0 : tg / S;
  const int s = pair_mode ? row : tg - rl * S;
  const int kb = (rl < RPG ? rl : 0) * S;
  const int ntiles = (a.R + RPG - 1) / RPG;
  const int tile0 = pair_mode ? blockIdx.x * (NG / 2) + pair : blockIdx.x * NG + grp;
  const int tstep = pair_mode ? gridDim.x * (NG / 2) : gridDim.x * NG;

  for (int tile = tile0; tile < ntiles; tile += tstep) {
    const int ray = tile * RPG + rl;
    const bool act = (rl < RPG) && (s < S) && (ray < a.R);
    const size_t smp = act ? ((size_t)ray * S + s) : 0;
    const float4* sp = reinterpret_cast<const float4*>(a.stash) + (size_t)(pair_mode ? 2 * tile + gp : tile) * (RP_PLANES * GROUP) + tg;

    // K / V / Q rows of this sample: global -> shared, asynchronously (rows of inactive samples are zero-filled)
    if (act) {
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        cp_async16(sq + row * 16 + 4 * d, sp + (RP_Q + d) * GROUP);
        cp_async16(sk + row * 16 + 4 * d, sp + (RP_K + d) * GROUP);
        cp_async16(sv + row * 16 + 4 * d, sp + (RP_V + d) * GROUP);
      }
    } else {
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        *reinterpret_cast<float4*>(sq + row * 16 + 4 * d) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(sk + row * 16 + 4 * d) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(sv + row * 16 + 4 * d) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // next tile of this group towards L2
    {
      const int nt = tile + tstep;
      if (nt < ntiles && (tg & 7) == 0) {
        const float4* np = reinterpret_cast<const float4*>(a.stash) + (size_t)(pair_mode ? 2 * nt + gp : nt) * (RP_PLANES * GROUP) + tg;
#pragma unroll 5
        for (int pl = 0; pl < RP_PLANES; ++pl) asm volatile("prefetch.global.L2 [%0];" ::"l"(np + pl * GROUP));
      }
    }

    float4 misc = make_float4(1.f, 0.f, 0.f, 0.f);
    float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t hq[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    float xhat[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) xhat[c] = 0.f;
    if (act) {
      misc = __ldcs(sp + RP_MISC * GROUP);
      dr = __ldg(reinterpret_cast<const float4*>(a.d_raw) + smp);
      rp_ld_codes8(sp, RP_HH, hq);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 q = __ldcs(sp + (RP_XHAT + j) * GROUP);
        xhat[4 * j] = q.x; xhat[4 * j + 1] = q.y; xhat[4 * j + 2] = q.z; xhat[4 * j + 3] = q.w;
      }
    }
    const float rstd = misc.x, z2 = misc.y, nvalid = misc.z;
    const bool row_valid = act && nvalid > 1.f;
    const float dz2 = (z2 > 0.f && !(nvalid < 1.f)) ? dr.w : 0.f;
    // sigma head
    {
      float dh[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dh[2 * j] = dz2 * sf[RF_W_OG2 + 2 * j] * elu_stash_lo(hq[j]);
        dh[2 * j + 1] = dz2 * sf[RF_W_OG2 + 2 * j + 1] * elu_stash_hi(hq[j]);
      }
      r_put16<NPASS>(tl, 0, dh);
    }
    NFB_RTC_BWD(RL_OG0);
    NFB_RTC_WAIT();
    // LayerNorm backward: dy = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dln * gamma
    {
      float dy[16];
      r_ld16(tl, 0, dy);
      float gsum = 0.f, gx = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        dy[c] *= sf[RF_LNW + c];
        gsum += dy[c];
        gx = fmaf(dy[c], xhat[c], gx);
      }
      gsum *= (1.f / 16.f); gx *= (1.f / 16.f);
      uint32_t park[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        dy[c] = rstd * (dy[c] - gsum - xhat[c] * gx);
        park[c] = __float_as_uint(dy[c]);
      }
      tmem_st16(tl + RC_D + 48, park);          // dy is needed again after the attention passes
      r_put16<NPASS>(tl, 0, dy);
    }
    NFB_RTC_BWD(RL_FC);
    NFB_RTC_WAIT();
    float2 dO01[4], dO23[4];
    float nm[4], il[4], nD[4];
    {
      float dO[16];
      r_ld16(tl, 0, dO);
      float o[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) o[c] = 0.f;
      float4 nm4 = make_float4(0.f, 0.f, 0.f, 0.f), il4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (act) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 q = __ldcs(sp + (RP_O + j) * GROUP);
          o[4 * j] = q.x; o[4 * j + 1] = q.y; o[4 * j + 2] = q.z; o[4 * j + 3] = q.w;
        }
        nm4 = __ldcs(sp + RP_NM * GROUP);
        il4 = __ldcs(sp + RP_IL * GROUP);
      }
      nm[0] = nm4.x; nm[1] = nm4.y; nm[2] = nm4.z; nm[3] = nm4.w;
      il[0] = il4.x; il[1] = il4.y; il[2] = il4.z; il[3] = il4.w;
#pragma unroll
      for (int h = 0; h < 4; ++h)
        nD[h] = -(dO[4 * h] * o[4 * h] + dO[4 * h + 1] * o[4 * h + 1] + dO[4 * h + 2] * o[4 * h + 2] + dO[4 * h + 3] * o[4 * h + 3]);
      const float vf = row_valid ? 1.f : 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d)
        *reinterpret_cast<float4*>(sdo + row * 16 + 4 * d) = make_float4(dO[d], dO[4 + d], dO[8 + d], dO[12 + d]);
      *reinterpret_cast<float4*>(sst + row * 16) = nm4;
      *reinterpret_cast<float4*>(sst + row * 16 + 4) = il4;
      *reinterpret_cast<float4*>(sst + row * 16 + 8) = make_float4(nD[0], nD[1], nD[2], nD[3]);
      *reinterpret_cast<float4*>(sst + row * 16 + 12) = make_float4(vf, vf, vf, vf);
    }
    cp_async_wait_all();
    named_bar_sync(att_bar, att_n);

    // query side: dq_i = sum_j dS_ij k_j  (zero for masked rows: masked_fill blocks the gradient)
    {
      float2 q01[4], q23[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const float4 q4 = *reinterpret_cast<const float4*>(sq + row * 16 + 4 * d);
        q01[d] = make_float2(q4.x, q4.y);
        q23[d] = make_float2(q4.z, q4.w);
        // dO pairs re-read from the row just published: LDS.128 delivers (head 0, head 1) / (head 2, head 3) in ALIGNED register
        // pairs.  Built from the TMEM read-back (head-major registers) the compiler re-packed all eight pairs with IMAD.MOV
        // in every iteration of the loop below (11.5 % of the kernel's instructions, ncu r02d).
        const float4 g4 = *reinterpret_cast<const float4*>(sdo + row * 16 + 4 * d);
        dO01[d] = make_float2(g4.x, g4.y);
        dO23[d] = make_float2(g4.z, g4.w);
      }
      const float4* kr = reinterpret_cast<const float4*>(sk + kb * 16);
      const float4* vr = reinterpret_cast<const float4*>(sv + kb * 16);
      const float2 nm01 = make_float2(nm[0], nm[1]), nm23 = make_float2(nm[2], nm[3]);
      const float2 il01 = make_float2(il[0], il[1]), il23 = make_float2(il[2], il[3]);
      const float2 nD01 = make_float2(nD[0], nD[1]), nD23 = make_float2(nD[2], nD[3]);
      float2 dq01[4], dq23[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) dq01[d] = dq23[d] = make_float2(0.f, 0.f);
      for (int j = 0; j < S; ++j) {
        float2 s01, s23;
        attn_scores(kr + 4 * j, q01, q23, s01, s23);
        s01 = __fadd2_rn(s01, nm01);
        s23 = __fadd2_rn(s23, nm23);
        const float2 p01 = __fmul2_rn(make_float2(ex2_approx(s01.x), ex2_approx(s01.y)), il01);
        const float2 p23 = __fmul2_rn(make_float2(ex2_approx(s23.x), ex2_approx(s23.y)), il23);
        float2 dP01 = nD01, dP23 = nD23;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const float4 vj = vr[4 * j + d];
          dP01 = __ffma2_rn(dO01[d], make_float2(vj.x, vj.y), dP01);
          dP23 = __ffma2_rn(dO23[d], make_float2(vj.z, vj.w), dP23);
        }
        const float2 dS01 = __fmul2_rn(p01, dP01), dS23 = __fmul2_rn(p23, dP23);
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const float4 kj = kr[4 * j + d];
          dq01[d] = __ffma2_rn(dS01, make_float2(kj.x, kj.y), dq01[d]);
          dq23[d] = __ffma2_rn(dS23, make_float2(kj.z, kj.w), dq23[d]);
        }
      }
      const float sc = row_valid ? INV_TEMP : 0.f;
      float dq[16];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        dq[d] = dq01[d].x * sc;
        dq[4 + d] = dq01[d].y * sc;
        dq[8 + d] = dq23[d].x * sc;
        dq[12 + d] = dq23[d].y * sc;
      }
      r_put16<NPASS>(tl, 0, dq);
    }
    // key side: dk_j = sum_i dS_ij q_i ; dv_j = sum_i p_ij dO_i   (this thread is key j)
    {
      float2 k01[4], k23[4], v01[4], v23[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const float4 k4 = *reinterpret_cast<const float4*>(sk + row * 16 + 4 * d);
        const float4 v4 = *reinterpret_cast<const float4*>(sv + row * 16 + 4 * d);
        k01[d] = make_float2(k4.x, k4.y); k23[d] = make_float2(k4.z, k4.w);
        v01[d] = make_float2(v4.x, v4.y); v23[d] = make_float2(v4.z, v4.w);
      }
      float2 dk01[4], dk23[4], dv01[4], dv23[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) dk01[d] = dk23[d] = dv01[d] = dv23[d] = make_float2(0.f, 0.f);
      const float4* qr = reinterpret_cast<const float4*>(sq + kb * 16);
      const float4* gr = reinterpret_cast<const float4*>(sdo + kb * 16);
      const float4* st = reinterpret_cast<const float4*>(sst + kb * 16);
      for (int i = 0; i < S; ++i) {
        float4 qi[4], gi[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) { qi[d] = qr[4 * i + d]; gi[d] = gr[4 * i + d]; }
        const float4 nmi = st[4 * i], ili = st[4 * i + 1], nDi = st[4 * i + 2], vf = st[4 * i + 3];
        float2 s01 = make_float2(nmi.x, nmi.y), s23 = make_float2(nmi.z, nmi.w);
        float2 dP01 = make_float2(nDi.x, nDi.y), dP23 = make_float2(nDi.z, nDi.w);
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          s01 = __ffma2_rn(make_float2(qi[d].x, qi[d].y), k01[d], s01);
          s23 = __ffma2_rn(make_float2(qi[d].z, qi[d].w), k23[d], s23);
          dP01 = __ffma2_rn(make_float2(gi[d].x, gi[d].y), v01[d], dP01);
          dP23 = __ffma2_rn(make_float2(gi[d].z, gi[d].w), v23[d], dP23);
        }
        const float2 p01 = __fmul2_rn(make_float2(ex2_approx(s01.x), ex2_approx(s01.y)), make_float2(ili.x, ili.y));
        const float2 p23 = __fmul2_rn(make_float2(ex2_approx(s23.x), ex2_approx(s23.y)), make_float2(ili.z, ili.w));
        const float2 dS01 = __fmul2_rn(__fmul2_rn(p01, dP01), make_float2(vf.x, vf.y));
        const float2 dS23 = __fmul2_rn(__fmul2_rn(p23, dP23), make_float2(vf.z, vf.w));
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          dk01[d] = __ffma2_rn(dS01, make_float2(qi[d].x, qi[d].y), dk01[d]);
          dk23[d] = __ffma2_rn(dS23, make_float2(qi[d].z, qi[d].w), dk23[d]);
          dv01[d] = __ffma2_rn(p01, make_float2(gi[d].x, gi[d].y), dv01[d]);
          dv23[d] = __ffma2_rn(p23, make_float2(gi[d].z, gi[d].w), dv23[d]);
        }
      }
      float t[16];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        t[d] = dk01[d].x * (1.f / LOG2E);
        t[4 + d] = dk01[d].y * (1.f / LOG2E);
        t[8 + d] = dk23[d].x * (1.f / LOG2E);
        t[12 + d] = dk23[d].y * (1.f / LOG2E);
      }
      r_put16<NPASS>(tl, 1, t);
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        t[d] = dv01[d].x; t[4 + d] = dv01[d].y; t[8 + d] = dv23[d].x; t[12 + d] = dv23[d].y;
      }
      r_put16<NPASS>(tl, 2, t);
    }
    // d xin = dy (residual) + [dq | dk | dv] Wqkv ; pos_encoding is a constant
    NFB_RTC_BWD(RL_QKV);
    uint32_t gq[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (act) rp_ld_codes8(sp, RP_G16, gq);
    NFB_RTC_WAIT();
    {
      float dx[16];
      r_ld16(tl, 0, dx);
      uint32_t park[16];
      tmem_ld16u(tl + RC_D + 48, park);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dx[2 * j] = (dx[2 * j] + __uint_as_float(park[2 * j])) * elu_stash_lo(gq[j]);
        dx[2 * j + 1] = (dx[2 * j + 1] + __uint_as_float(park[2 * j + 1])) * elu_stash_hi(gq[j]);
      }
      r_put16<NPASS>(tl, 0, dx);
    }
    NFB_RTC_BWD(RL_GEO2);
    uint32_t cq[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) cq[j] = 0u;
    if (act) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 q = __ldcs(sp + (RP_H64 + j) * GROUP);
        cq[4 * j] = __float_as_uint(q.x); cq[4 * j + 1] = __float_as_uint(q.y);
        cq[4 * j + 2] = __float_as_uint(q.z); cq[4 * j + 3] = __float_as_uint(q.w);
      }
    }
    NFB_RTC_WAIT();
    float dwm = 0.f;
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      float dh[16];
      r_ld16(tl, 16 * kc, dh);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dh[2 * j] *= elu_stash_lo(cq[8 * kc + j]);
        dh[2 * j + 1] *= elu_stash_hi(cq[8 * kc + j]);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) dwm = fmaf(dh[j], sf[RF_WCOL + 16 * kc + j], dwm);
      r_put16<NPASS>(tl, kc, dh);
    }
    NFB_RTC_BWD(RL_GEO0);
    NFB_RTC_WAIT();
    {
      float4* out = reinterpret_cast<float4*>(a.d_ps + smp * NFB_PS_STRIDE);
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        float t[16];
        r_ld16(tl, 16 * kc, t);
        if (act) {
#pragma unroll
          for (int j = 0; j < 4; ++j) out[4 * kc + j] = make_float4(t[4 * j], t[4 * j + 1], t[4 * j + 2], t[4 * j + 3]);
        }
      }
      if (act) {
        out[16] = make_float4(dwm, dr.x, dr.y, dr.z);
        out[17] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    named_bar_sync(att_bar, att_n);                    // K / V / Q / dO rows are rewritten by the next tile
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*s_tmem, NG * 128);
}

template <int NPASS>
int launch_ray_tc_bwd_stash(const RayArgs& a, cudaStream_t st) {
  const size_t smem = bwd_stash_smem<NPASS>();
  cudaError_t e = cudaFuncSetAttribute(k_ray_tc_bwd_stash<NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "k_ray_tc_bwd_stash: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const bool pair_mode = a.S > GROUP;
  const int RPG = pair_mode ? 1 : GROUP / a.S;
  const int ntiles = (a.R + RPG - 1) / RPG;
  const int per_cta = pair_mode ? BS_NG / 2 : BS_NG;
  int grid = (ntiles + per_cta - 1) / per_cta;
  const int cap = nfb_num_sms();
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  k_ray_tc_bwd_stash<NPASS><<<grid, GROUP * BS_NG, smem, st>>>(a);
  NFB_CHECK_LAUNCH("k_ray_tc_bwd_stash");
  return NFB_OK;
}

}  // namespace nfbrtc

// defined in nfb_ray_tc_inst.cu (one instantiation per translation unit)
int nfb_launch_ray_tc_fwd_p1(const nfbrtc::RayArgs& a, cudaStream_t st);
int nfb_launch_ray_tc_fwd_p3(const nfbrtc::RayArgs& a, cudaStream_t st);
int nfb_launch_ray_tc_bwd_p1(const nfbrtc::RayArgs& a, cudaStream_t st);
int nfb_launch_ray_tc_bwd_p3(const nfbrtc::RayArgs& a, cudaStream_t st);
int nfb_launch_ray_tc_fwd_p1_save(const nfbrtc::RayArgs& a, cudaStream_t st);
int nfb_launch_ray_tc_fwd_p3_save(const nfbrtc::RayArgs& a, cudaStream_t st);
int nfb_launch_ray_tc_bwd_stash_p1(const nfbrtc::RayArgs& a, cudaStream_t st);
int nfb_launch_ray_tc_bwd_stash_p3(const nfbrtc::RayArgs& a, cudaStream_t st);
