// IBRNet view stage on the 5th-generation tensor cores (tcgen05 + TMEM), forward.
//
// Same mathematics and row mapping as the fp32 CUDA-core form in nfb_view_stage.cuh -- one thread per
// (sample, view) row, 128-row groups, cross-view reductions through a per-group exchange buffer -- but every
// dense layer with >= 16 inputs runs as a 128 x N x K tcgen05.mma tile:
//   * the thread that owns row i writes the layer input of its row as bf16 into TMEM lane i (tcgen05.st),
//   * one thread of the group issues the K/16 MMAs against the layer's weight tile resident in shared memory
//     (canonical no-swizzle K-major layout, nfb_tc.cuh) and commits them to the group's mbarrier,
//   * every thread reads its own row of the fp32 accumulator back (tcgen05.ld) and applies bias + ELU.
// NG = 4 groups per CTA (one CTA per SM, 512 TMEM columns = 4 x 128) keep the CUDA cores busy while a group
// waits for its MMAs.
//
// Precision (NPASS): 1 = plain bf16 operands (fp32 accumulate).  3 = "bf16x3": activations and weights are
// split x = hi + lo (two bf16, ~17 significant bits) and D = hi*hi + lo*hi + hi*lo is accumulated in fp32 --
// products are exact to ~2^-17, which keeps the whole path inside the reference's fp32 tolerance.
//
// TMEM columns of a group (128): D [0,64)  |  1-pass: A [64,120)   3-pass: A_hi [64,96), A_lo [96,128).
// With 3 passes the 112-wide input of base_fc.0 does not fit in 32+32 columns, so that layer is issued in
// two rounds (K = 64, then K = 48) accumulating into the same D.
#pragma once
#include "nfb_dense.cuh"
#include "nfb_geom.cuh"
#include "nfb_tc.cuh"
#include "nfb_view_stage.cuh"

namespace nfbvtc {
using namespace nfbtc;
using nfbview::ViewArgs;

constexpr int GROUP = 128;
#ifndef NFB_VTC_NG
#define NFB_VTC_NG 4          // 128-row groups per CTA (4 x 128 TMEM columns = the whole TMEM of the SM)
#endif
constexpr int NG = NFB_VTC_NG;
constexpr int TMEM_ALLOC = NG * 128 > 256 ? 512 : (NG * 128 > 128 ? 256 : 128);   // tcgen05.alloc takes powers of two
constexpr int EXS = 37;      // exchange-buffer row stride (floats)
constexpr int TS_MAX = 32;   // samples per tile cap (V < 4 leaves part of a 128-row tile idle)
constexpr int MVS = 72;      // per-sample pooled statistics: mean0[35] at 0, var0[35] at 36
constexpr int MVP = 72;      // 32-bit words per sample of the packed copy: bf16 hi halves of [mean | var] (70 values =
                             // 35 words) at 0, lo halves at 36 -- already in base_fc.0 operand order
constexpr int GC = 128;      // TMEM columns per group
constexpr int C_D = 0, C_A = 64, C_ALO = 96;
constexpr int EXQ = 36;      // exchange-buffer row stride of the forward / stash-backward kernels (floats): 144 B rows are 16-byte
                             // aligned and 8 consecutive rows start in 8 different 16-byte bank groups -> the float4 loads /
                             // stores of the cross-view poolings are conflict-free
constexpr int SIDE = 4;      // per-row side slots (rgb_in[3] for the blending, 1 spare)

// ---- thread -> (sample, view) mapping of a 128-row tile ------------------------------------------------------------
// contiguous: row r of the tile = thread r = (sample r / V, view r % V); a sample's rows straddle warps unless V | 32.
// packed:     every warp holds spw = floor(32 / V) whole samples in its first spw * V lanes (the rest idle), so the V rows
//             of a sample always share a warp: the cross-view exchanges need __syncwarp() only and may use shuffles.
//             Chosen when it keeps >= 90 % of the samples per tile of the contiguous form (V = 10: 12 samples either way).
struct RowMap {
  int TS;        // samples per tile
  int spw;       // samples per warp (packed form)
  bool packed;
};
__host__ __device__ inline RowMap row_map(int V) {
  RowMap m;
  const int ts_c = (GROUP / V < TS_MAX) ? GROUP / V : TS_MAX;
  int spw = V <= 32 ? 32 / V : 0;
  if (spw > TS_MAX / 4) spw = TS_MAX / 4;
  const int ts_w = 4 * spw;
  m.packed = spw >= 1 && ts_w * 10 >= ts_c * 9;
  m.TS = m.packed ? ts_w : ts_c;
  m.spw = spw;
  return m;
}
// exchange-buffer slot (= thread index within the group) of row rr = sample * V + view of the tile
__device__ __forceinline__ int row_slot(const RowMap& m, int V, int rr) {
  if (!m.packed) return rr;
  const int s = rr / V, vv = rr - s * V;
  return (s / m.spw) * 32 + (s % m.spw) * V + vv;
}

// ---- weight tiles ---------------------------------------------------------------------------------
enum : int { L_DIR2 = 0, L_BASE0, L_BASE2, L_VIS0, L_VIS2, L_VISB0, L_RGB0, L_COUNT };
__host__ __device__ constexpr int layer_n(int l) {
  return l == L_DIR2 ? 48 : l == L_BASE0 ? 64 : l == L_BASE2 ? 32 : l == L_VIS0 ? 32 : l == L_VIS2 ? 48 : l == L_VISB0 ? 32 : 16;
}
__host__ __device__ constexpr int layer_k(int l) {
  return l == L_DIR2 ? 16 : l == L_BASE0 ? 112 : l == L_BASE2 ? 64 : l == L_VIS0 ? 32 : l == L_VIS2 ? 32 : l == L_VISB0 ? 32 : 48;
}
__host__ __device__ constexpr int layer_off(int l) {   // byte offset of the layer's tile inside one (hi or lo) set
  int o = 0;
  for (int i = 0; i < l; ++i) o += layer_n(i) * layer_k(i) * 2;
  return o;
}
constexpr int B_SET_BYTES = layer_off(L_COUNT);   // 28672
static_assert(B_SET_BYTES % 16 == 0, "tile alignment");

// canonical K-major no-swizzle byte offset of element (n, k) in an [N][K] bf16 tile stored K-chunk-major
__host__ __device__ constexpr uint32_t canon_off(int n, int k, int N) {
  return (uint32_t)((k >> 3) * (N * 16) + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2);
}

// ---- fp32 side tables (small layers that stay on the CUDA cores, biases) --------------------------
enum : int {
  F_DIR0_W = 0,                 // [4][16]  ray_dir_fc.0 transposed
  F_DIR0_B = F_DIR0_W + 64,     // 16
  F_B_DIR2 = F_DIR0_B + 16,     // 48
  F_B_BASE0 = F_B_DIR2 + 48,    // 64
  F_B_BASE2 = F_B_BASE0 + 64,   // 32
  F_B_VIS0 = F_B_BASE2 + 32,    // 32
  F_B_VIS2 = F_B_VIS0 + 32,     // 48
  F_B_VISB0 = F_B_VIS2 + 48,    // 32
  F_W_VISB2 = F_B_VISB0 + 32,   // 32
  F_B_VISB2 = F_W_VISB2 + 32,   // 4
  F_B_RGB0 = F_B_VISB2 + 4,     // 16
  F_W_RGB2 = F_B_RGB0 + 16,     // [16][8] transposed
  F_B_RGB2 = F_W_RGB2 + 128,    // 8
  F_W_RGB4 = F_B_RGB2 + 8,      // 8
  F_B_RGB4 = F_W_RGB4 + 8,      // 4
  F_S = F_B_RGB4 + 4,           // 4
  F_TOTAL = F_S + 4
};

template <int NPASS>
__host__ __device__ constexpr size_t smem_bytes() {
  return (size_t)B_SET_BYTES * (NPASS == 3 ? 2 : 1) +
         sizeof(float) * (F_TOTAL + 16 * NFB_MAX_VIEWS + 4 + NG * GROUP * (EXQ + SIDE) + NG * TS_MAX * MVP) + NG * 8 + 16;
}

template <int NPASS>
static __device__ void load_tile(uint8_t* sB, int layer, const float* __restrict__ w, int n_real, int k_real, int tid, int nt) {
  const int N = layer_n(layer), K = layer_k(layer);
  uint8_t* hi = sB + layer_off(layer);
  uint8_t* lo = hi + B_SET_BYTES;
  for (int i = tid; i < N * K; i += nt) {
    const int n = i / K, k = i - n * K;
    const float v = (n < n_real && k < k_real) ? __ldg(w + n * k_real + k) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const uint32_t off = canon_off(n, k, N);
    *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
    if (NPASS == 3) *reinterpret_cast<__nv_bfloat16*>(lo + off) = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

static __device__ void load_side_tables(float* sf, const float* __restrict__ p, int tid, int nt) {
  load_wt_transposed(sf + F_DIR0_W, p + P_DIR0_W, 16, 4, 16, tid, nt);
  load_vec_padded(sf + F_DIR0_B, p + P_DIR0_B, 16, 16, tid, nt);
  load_vec_padded(sf + F_B_DIR2, p + P_DIR2_B, 35, 48, tid, nt);
  load_vec_padded(sf + F_B_BASE0, p + P_BASE0_B, 64, 64, tid, nt);
  load_vec_padded(sf + F_B_BASE2, p + P_BASE2_B, 32, 32, tid, nt);
  load_vec_padded(sf + F_B_VIS0, p + P_VIS0_B, 32, 32, tid, nt);
  load_vec_padded(sf + F_B_VIS2, p + P_VIS2_B, 33, 48, tid, nt);
  load_vec_padded(sf + F_B_VISB0, p + P_VISB0_B, 32, 32, tid, nt);
  load_vec_padded(sf + F_W_VISB2, p + P_VISB2_W, 32, 32, tid, nt);
  load_vec_padded(sf + F_B_VISB2, p + P_VISB2_B, 1, 4, tid, nt);
  load_vec_padded(sf + F_B_RGB0, p + P_RGB0_B, 16, 16, tid, nt);
  load_wt_transposed(sf + F_W_RGB2, p + P_RGB2_W, 8, 16, 8, tid, nt);
  load_vec_padded(sf + F_B_RGB2, p + P_RGB2_B, 8, 8, tid, nt);
  load_vec_padded(sf + F_W_RGB4, p + P_RGB4_W, 8, 8, tid, nt);
  load_vec_padded(sf + F_B_RGB4, p + P_RGB4_B, 1, 4, tid, nt);
  if (tid == 0) sf[F_S] = fabsf(__ldg(p + P_S));
}

// ---- per-thread TMEM helpers ------------------------------------------------------------------------
// write 16 consecutive layer inputs (K-chunk `kc` of the A operand) of this thread's row
template <int NPASS>
__device__ __forceinline__ void a_store16(uint32_t tl, int kc, const float (&v)[16]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (NPASS == 3) split_bf16(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
    else hi[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
  }
  tmem_st8(tl + C_A + 8 * kc, hi);
  if (NPASS == 3) tmem_st8(tl + C_ALO + 8 * kc, lo);
}

// same, from 8 already packed words (hi) and their lo twins
template <int NPASS>
__device__ __forceinline__ void a_store_words(uint32_t tl, int kc, const uint32_t (&hi)[8], const uint32_t (&lo)[8]) {
  tmem_st8(tl + C_A + 8 * kc, hi);
  if (NPASS == 3) tmem_st8(tl + C_ALO + 8 * kc, lo);
}
template <int NPASS>
__device__ __forceinline__ void pack_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (NPASS == 3) split_bf16(a, b, hi, lo);
  else { hi = pack_bf16(a, b); lo = 0u; }
}

// y[j] = ELU(D[col + j] + bias[j]), 16 outputs
__device__ __forceinline__ void epi16(uint32_t tl, int col, const float* __restrict__ bias, float (&y)[16]) {
  tmem_ld16(tl + C_D + col, y);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias + j);
    const float2 r0 = elu_fast2(__fadd2_rn(make_float2(y[j + 0], y[j + 1]), make_float2(b.x, b.y)));
    const float2 r1 = elu_fast2(__fadd2_rn(make_float2(y[j + 2], y[j + 3]), make_float2(b.z, b.w)));
    y[j + 0] = r0.x; y[j + 1] = r0.y; y[j + 2] = r1.x; y[j + 3] = r1.y;
  }
}

// MMAs of k-steps [KS0, KS1) of `layer`; the A operand of k-step ks sits in A chunk (ks - KS0).  One thread.
template <int NPASS, int LAYER, int KS0, int KS1>
__device__ __forceinline__ void issue_mma(uint32_t tb, uint32_t sB_addr, bool acc0) {
  constexpr int N = layer_n(LAYER);
  constexpr uint32_t idesc = idesc_bf16(128, N);
  const uint32_t bhi = sB_addr + layer_off(LAYER), blo = bhi + B_SET_BYTES;
#pragma unroll
  for (int ks = KS0; ks < KS1; ++ks) {
    const uint64_t dh = smem_desc(bhi + ks * 2 * N * 16, N * 16, 128);
    const uint32_t ah = tb + C_A + 8 * (ks - KS0);
    mma_ts(tb + C_D, ah, dh, idesc, acc0 || ks > KS0);
    if (NPASS == 3) {
      const uint64_t dl = smem_desc(blo + ks * 2 * N * 16, N * 16, 128);
      mma_ts(tb + C_D, tb + C_ALO + 8 * (ks - KS0), dh, idesc, true);
      mma_ts(tb + C_D, ah, dl, idesc, true);
    }
  }
}

// element k of the base_fc.0 input row: mean0[k] (k < 35), var0[k - 35] (k < 70), x0[k - 70] (k < 105), 0 (pad);
// k is a compile-time constant after unrolling, so the branches fold and x stays in registers
__device__ __forceinline__ float base0_in(int k, const float* __restrict__ mvs, const float (&x)[NFB_ROW_CH]) {
  if (k < 35) return mvs[k];
  if (k < 70) return mvs[36 + (k - 35)];
  if (k < 105) return x[k - 70];
  return 0.f;
}

// ---- activation stash (forward with SAVE -> nfb_view_tc_bwd2.cuh) ----------------------------------------
// What the data-gradient needs from the forward, written once instead of recomputed: per 128-row tile
// ST_PLANES planes of [128 rows][4 words] (a thread's 16-byte vector sits next to its neighbours': coalesced).
//   planes  0..8   x0 = rgb_feat + direction_feat (35 fp32, 1 pad)
//   planes  9..16  x2 (32 fp32)
//   plane   17     w, mask, sigmoid(vis_fc.2[32]), sigmoid(vis_fc2)
//   plane   18     blending logit, grid x, grid y, rgb_in[0]
//   plane   19     rgb_in[1], rgb_in[2], ELU' code of vis_fc.2[32], pad
//   planes 20..27  ELU' codes of h1 (base_fc.0 out, 64)      28..31 x1 (base_fc.2 out, 32)
//   planes 32..35  hv (vis_fc.0 out, 32)   36..39 xv[0..32) (vis_fc.2 out)   40..43 hv2 (vis_fc2.0 out, 32)
//   planes 44..45  g1 (rgb_fc.0 out, 16)   46 g2 (rgb_fc.2 out, 8)   47 spare
enum : int { SP_X0 = 0, SP_X2 = 9, SP_SA = 17, SP_SB = 18, SP_SC = 19, SP_H1 = 20, SP_X1 = 28, SP_HV = 32, SP_XV = 36,
             SP_HV2 = 40, SP_G1 = 44, SP_G2 = 46, ST_PLANES = 48 };
constexpr size_t ST_TILE_BYTES = (size_t)ST_PLANES * GROUP * 16;   // 98304

// ELU and the 16-bit code of its derivative from one exponential: y = max(x, min(e - 1, 0)), ELU' = min(e, 1) = 2 - t
// with t = max(1.9999999 - e, 1) in [1, 2) (nfb_tc.cuh: elu_stash_*)
__device__ __forceinline__ void elu_with_t(float x, float& y, float& t) {
  const float e = ex2_approx(x * 1.4426950408889634f);
  y = fmaxf(x, fminf(e - 1.f, 0.f));
  t = fmaxf(1.9999999f - e, 1.f);
}
__device__ __forceinline__ uint32_t pack_t(float t0, float t1) {
  return __byte_perm(__float_as_uint(t0), __float_as_uint(t1), 0x6521);
}
// y[j] = ELU(D[col + j] + bias[j]) and the derivative codes of the 16 outputs
__device__ __forceinline__ void epi16_code(uint32_t tl, int col, const float* __restrict__ bias, float (&y)[16], uint32_t (&q)[8]) {
  tmem_ld16(tl + C_D + col, y);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias + j);
    const float2 r0 = elu_code2(__fadd2_rn(make_float2(y[j + 0], y[j + 1]), make_float2(b.x, b.y)), q[j / 2]);
    const float2 r1 = elu_code2(__fadd2_rn(make_float2(y[j + 2], y[j + 3]), make_float2(b.z, b.w)), q[j / 2 + 1]);
    y[j + 0] = r0.x; y[j + 1] = r0.y; y[j + 2] = r1.x; y[j + 3] = r1.y;
  }
}
template <bool SAVE>
__device__ __forceinline__ void epi16_s(uint32_t tl, int col, const float* __restrict__ bias, float (&y)[16], uint32_t (&q)[8]) {
  if (SAVE) epi16_code(tl, col, bias, y, q);
  else epi16(tl, col, bias, y);
}
// Two 16-column chunks with ONE tcgen05.wait::ld (the asm statements are volatile: two epi16 calls in a row serialise
// load -> wait -> compute -> load -> wait; here both loads are in flight before the wait).  NFB_VTC_PAIR selects it.
#ifndef NFB_VTC_PAIR
#define NFB_VTC_PAIR 0
#endif
#ifndef NFB_VTC_COOP_GATHER
#define NFB_VTC_COOP_GATHER 1     // quarter-warp cooperative feature gather (gather_row_coop, nfb_geom.cuh); 0 = one lane per row (gather_row): 140.9 vs 120.9 ms
#endif
template <bool SAVE>
__device__ __forceinline__ void epi16_pair(uint32_t tl, int col, const float* __restrict__ bias, float (&y0)[16], float (&y1)[16],
                                           uint32_t (&q0)[8], uint32_t (&q1)[8]) {
  tmem_ld16(tl + C_D + col, y0);
  tmem_ld16(tl + C_D + col + 16, y1);
  tmem_ld_wait();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float (&y)[16] = h ? y1 : y0;
    uint32_t (&q)[8] = h ? q1 : q0;
    const float* b_ = bias + 16 * h;
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(b_ + j);
      const float2 s0 = __fadd2_rn(make_float2(y[j + 0], y[j + 1]), make_float2(b.x, b.y));
      const float2 s1 = __fadd2_rn(make_float2(y[j + 2], y[j + 3]), make_float2(b.z, b.w));
      float2 r0, r1;
      if (SAVE) { r0 = elu_code2(s0, q[j / 2]); r1 = elu_code2(s1, q[j / 2 + 1]); }
      else { r0 = elu_fast2(s0); r1 = elu_fast2(s1); }
      y[j + 0] = r0.x; y[j + 1] = r0.y; y[j + 2] = r1.x; y[j + 3] = r1.y;
    }
  }
}
__device__ __forceinline__ void st_plane(float4* sp, int plane, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  __stcs(sp + plane * GROUP, make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c), __uint_as_float(d)));   // streaming: written once, read once by the backward
}
__device__ __forceinline__ void st_codes8(float4* sp, int plane, const uint32_t (&q)[8]) {
  st_plane(sp, plane, q[0], q[1], q[2], q[3]);
  st_plane(sp, plane + 1, q[4], q[5], q[6], q[7]);
}

// ---- cross-view reductions with instruction-level parallelism ---------------------------------------------
// The V rows of a sample sit in consecutive exchange-buffer rows.  A thread owns channels c0, c0 + V, c0 + 2V, ...;
// up to KMAX of them are accumulated side by side (independent chains, the loop over views outermost), so the
// shared-memory latency of one channel hides behind the others.  row0 = ex + base * EXS.
constexpr int POOL_K = 9;     // ceil(35 / V) for V >= 4; smaller V take several rounds
// acc[k] = sum_u row_u[c0 + k V] * (WEIGHTED ? row_u[wslot] * scale : 1)
template <bool WEIGHTED>
__device__ __forceinline__ void pool_sum(const float* __restrict__ row0, int V, int c0, int cmax, int wslot, float scale,
                                         float (&acc)[POOL_K]) {
#pragma unroll
  for (int k = 0; k < POOL_K; ++k) acc[k] = 0.f;
  for (int u = 0; u < V; ++u) {
    const float* r = row0 + u * EXS;
    const float wu = WEIGHTED ? r[wslot] * scale : 1.f;
#pragma unroll
    for (int k = 0; k < POOL_K; ++k) {
      const int c = c0 + k * V;
      if (c < cmax) acc[k] = WEIGHTED ? fmaf(r[c], wu, acc[k]) : acc[k] + r[c];
    }
  }
}
// var[k] = sum_u (row_u[wslot] * scale) * (row_u[c] - mean[k])^2
__device__ __forceinline__ void pool_var(const float* __restrict__ row0, int V, int c0, int cmax, int wslot, float scale,
                                         const float (&mean)[POOL_K], float (&var)[POOL_K]) {
#pragma unroll
  for (int k = 0; k < POOL_K; ++k) var[k] = 0.f;
  for (int u = 0; u < V; ++u) {
    const float* r = row0 + u * EXS;
    const float wu = r[wslot] * scale;
#pragma unroll
    for (int k = 0; k < POOL_K; ++k) {
      const int c = c0 + k * V;
      if (c < cmax) {
        const float d = r[c] - mean[k];
        var[k] = fmaf(wu * d, d, var[k]);
      }
    }
  }
}

// ---- cross-view pooling on float4 quads (forward / stash-backward kernels, row stride EXQ) ---------------------------------
// The V rows of a sample sit in consecutive exchange rows (row0 = view 0).  Thread v of the sample owns the channel quads
// q = v, v + V, ... < NQ; per quad and view one LDS.128 + one LDS.32 (the weight) instead of four guarded scalar loads.
// Same operation order as pool_sum / pool_var: mean = fma(x_u, w_u, mean) over u, var = fma(w_u * d, d, var), d = x_u - mean.
enum : int { POOL_SUM = 0, POOL_MEAN = 1, POOL_MEAN_VAR = 2 };
template <int NQ, int MODE, typename EMIT>
__device__ __forceinline__ void pool4(const float* __restrict__ row0, int V, int v, int wslot, float scale, EMIT&& emit) {
  for (int q = v; q < NQ; q += V) {
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int u = 0; u < V; ++u) {
      const float* r = row0 + u * EXQ;
      const float4 x = *reinterpret_cast<const float4*>(r + 4 * q);
      if (MODE != POOL_SUM) {
        const float wu = r[wslot] * scale;
        const float2 w2 = make_float2(wu, wu);                       // packed fp32x2: same FMAs, half the issue slots
        const float2 a = __ffma2_rn(make_float2(x.x, x.y), w2, make_float2(m.x, m.y));
        const float2 b = __ffma2_rn(make_float2(x.z, x.w), w2, make_float2(m.z, m.w));
        m = make_float4(a.x, a.y, b.x, b.y);
      } else {
        m.x += x.x; m.y += x.y; m.z += x.z; m.w += x.w;
      }
    }
    float4 s2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == POOL_MEAN_VAR) {
#pragma unroll 2
      for (int u = 0; u < V; ++u) {
        const float* r = row0 + u * EXQ;
        const float4 x = *reinterpret_cast<const float4*>(r + 4 * q);
        const float wu = r[wslot] * scale;
        const float2 w2 = make_float2(wu, wu), neg1 = make_float2(-1.f, -1.f);
        const float2 d01 = __ffma2_rn(make_float2(m.x, m.y), neg1, make_float2(x.x, x.y));       // x - m (exact as a subtraction)
        const float2 d23 = __ffma2_rn(make_float2(m.z, m.w), neg1, make_float2(x.z, x.w));
        const float2 a = __ffma2_rn(__fmul2_rn(w2, d01), d01, make_float2(s2.x, s2.y));
        const float2 b = __ffma2_rn(__fmul2_rn(w2, d23), d23, make_float2(s2.z, s2.w));
        s2 = make_float4(a.x, a.y, b.x, b.y);
      }
    }
    emit(q, m, s2);
  }
}

// all threads of the group: publish the A stores, let thread 0 issue + commit
// ISSUER (0..3): which warp of the group issues this layer.  Issuing an MMA batch costs its thread ~15 instructions per
// tcgen05.mma (descriptors, R2UR); rotating the issuer over the four warps keeps them in step (measured: with one fixed issuer
// that warp runs ~15 % more instructions per tile than the other three, and they wait for it at every barrier).
#define NFB_TC_ISSUE(LAYER, KS0, KS1, ACC0, ISSUER)                           \
  do {                                                                        \
    tmem_st_wait();                                                           \
    fence_before_sync();                                                      \
    named_bar_sync(bar_id, GROUP);                                            \
    if (tg == 32 * (ISSUER)) {                                                \
      fence_after_sync();                                                     \
      issue_mma<NPASS, LAYER, KS0, KS1>(tb, sB_addr, ACC0);                   \
      mma_commit(mbar);                                                       \
    }                                                                         \
  } while (0)
// Exchange-buffer synchronisation.  When V divides 32 the V rows of a sample are lanes of ONE warp, so the cross-view
// exchanges only need a warp barrier; otherwise the whole 128-row group synchronises.
#define NFB_EX_SYNC()                                   \
  do {                                                  \
    if (warp_local) __syncwarp();                       \
    else named_bar_sync(bar_id, GROUP);                 \
  } while (0)
#define NFB_TC_WAIT()          \
  do {                         \
    mbar_wait(mbar, phase);    \
    phase ^= 1u;               \
    fence_after_sync();        \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// forward kernel.  FUSED: rows come from projection + bilinear gather; otherwise from the materialised
// Projector.compute tensors (rgb_feat / ray_diff / mask).
// ---------------------------------------------------------------------------------------------------
template <int NPASS, bool FUSED, bool SAVE>
__global__ void __launch_bounds__(GROUP * NG, 1) k_view_tc_fwd(ViewArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sB = smem_raw;
  float* sf = reinterpret_cast<float*>(smem_raw + (size_t)B_SET_BYTES * (NPASS == 3 ? 2 : 1));
  float* s_cam = sf + F_TOTAL;
  float* ex_all = s_cam + (16 * NFB_MAX_VIEWS + 4);
  float* side_all = ex_all + NG * GROUP * EXQ;
  uint32_t* mvp_all = reinterpret_cast<uint32_t*>(side_all + NG * GROUP * SIDE);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(mvp_all + NG * TS_MAX * MVP);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + NG);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int grp = tid / GROUP, tg = tid % GROUP;
  float* ex = ex_all + (size_t)grp * GROUP * EXQ;
  float* side = side_all + (size_t)grp * GROUP * SIDE;
  uint32_t* mvp = mvp_all + (size_t)grp * TS_MAX * MVP;
  const int bar_id = 1 + grp;
  uint64_t* mbar = s_bar + grp;

  if (warp == 0) tmem_alloc(s_tmem, TMEM_ALLOC);
  if (tid == 0) {
    for (int g = 0; g < NG; ++g) mbar_init(s_bar + g, 1);
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
  fence_proxy_async_smem();     // weight tiles were written through the generic proxy, tcgen05.mma reads them through the async proxy
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const uint32_t tb = *s_tmem + (uint32_t)(grp * GC);                 // group's column base (lane field 0)
  const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16);       // + this warp's lane quarter
  const uint32_t sB_addr = smem_u32(sB);
  uint32_t phase = 0;

  const int V = a.V;
  const RowMap rm = row_map(V);
  const bool warp_local = rm.packed;
  const int TS = rm.TS;
  int sl, v;
  bool lane_ok;
  if (rm.packed) {
    const int wq = tg >> 5, l = tg & 31, si = l / V;
    v = l - si * V; sl = wq * rm.spw + si; lane_ok = si < rm.spw;
  } else {
    sl = tg / V; v = tg - sl * V; lane_ok = sl < TS;
  }
  const int ntiles = (a.N + TS - 1) / TS;
  const float Wm1 = (float)a.W - 1.f, Hm1 = (float)a.H - 1.f;
  const float s_abs = sf[F_S];

  for (int tile = blockIdx.x * NG + grp; tile < ntiles; tile += gridDim.x * NG) {
    const int p = tile * TS + sl;
    const bool active = lane_ok && (p < a.N);
    // rows without a sample (idle lanes, tile tail) still execute the exchange reads; they are pointed at rows / slots of their
    // OWN warp so that the warp-level fences of the packed mapping cover every shared-memory read (racecheck-clean)
    const int base = active ? tg - v : (tg & ~31);
    uint32_t* mvps = mvp + (active ? sl : (rm.packed ? (tg >> 5) * rm.spw : 0)) * MVP;
    float4* sp = reinterpret_cast<float4*>(a.stash) + (size_t)tile * (ST_PLANES * GROUP) + tg;
    const bool save = SAVE && active;

    // ---------------- projection, ray_diff, bilinear gather ----------------
    float x[NFB_ROW_CH];
    float rd[4];
    float mk = 0.f;
    float ggx = 0.f, ggy = 0.f;
    if (FUSED) {
#if NFB_VTC_COOP_GATHER
      ViewGeom g;
      g.gx = g.gy = g.mask = 0.f; g.rd[0] = g.rd[1] = g.rd[2] = g.rd[3] = 0.f;
      if (active) {
        float X, Y, Z;
        load_point(a.pts, p, X, Y, Z);
        g = view_geometry(X, Y, Z, s_cam + 16 * v, s_cam + 16 * V, Wm1, Hm1);
      }
      // the exchange rows are free here (the previous tile ended with a fence): they stage the quarter-warp transposition
      gather_row_coop(active, g, v, a.H, a.W, a.fh, a.fw, a.imgs, a.feat, ex, EXQ, tg, x);
      rd[0] = g.rd[0]; rd[1] = g.rd[1]; rd[2] = g.rd[2]; rd[3] = g.rd[3];
      mk = g.mask; ggx = g.gx; ggy = g.gy;
#else
      if (active) {
        float X, Y, Z;
        load_point(a.pts, p, X, Y, Z);
        const ViewGeom g = view_geometry(X, Y, Z, s_cam + 16 * v, s_cam + 16 * V, Wm1, Hm1);
        gather_row(g, v, a.H, a.W, a.fh, a.fw, a.imgs, a.feat, x);
        rd[0] = g.rd[0]; rd[1] = g.rd[1]; rd[2] = g.rd[2]; rd[3] = g.rd[3];
        mk = g.mask; ggx = g.gx; ggy = g.gy;
      } else {
#pragma unroll
        for (int c = 0; c < NFB_ROW_CH; ++c) x[c] = 0.f;
        rd[0] = rd[1] = rd[2] = rd[3] = 0.f;
      }
#endif
    } else {
      // the tile's rows are contiguous in rgb_feat: coalesced copy through the exchange buffer
      const size_t row0 = (size_t)tile * TS * V;
      const size_t total = (size_t)a.N * V;
      const int rows_here = (int)((total - row0 < (size_t)(TS * V)) ? (total - row0) : (size_t)(TS * V));
      const float* src = a.rgb_feat + row0 * NFB_ROW_CH;
      for (int i = tg; i < rows_here * NFB_ROW_CH; i += GROUP) {
        const int rr = i / NFB_ROW_CH, cc = i - rr * NFB_ROW_CH;
        ex[row_slot(rm, V, rr) * EXQ + cc] = __ldg(src + i);
      }
      named_bar_sync(bar_id, GROUP);
#pragma unroll
      for (int c = 0; c < NFB_ROW_CH; ++c) x[c] = active ? ex[tg * EXQ + c] : 0.f;
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

    // ---------------- ray_dir_fc.0 on the CUDA cores (K = 4), ray_dir_fc.2 as MMA ----------------
    {
      float a1[16];
      load_bias<16>(a1, sf + F_DIR0_B);
      dense_acc<4, 16>(sf + F_DIR0_W, rd, a1);
#pragma unroll
      for (int j = 0; j < 16; ++j) a1[j] = elu_fast(a1[j]);
      a_store16<NPASS>(tl, 0, a1);
    }
    NFB_TC_ISSUE(L_DIR2, 0, 1, false, 0);

    // ---------------- pooling weights (overlaps the MMA) ----------------
    float w, n_valid;
    {
      const float e = a.anti_alias ? (float)exp((double)__fmul_rn(s_abs, __fsub_rn(rd[3], 1.f))) : 1.f;
      *reinterpret_cast<float2*>(ex + tg * EXQ) = make_float2(e, mk);
      NFB_EX_SYNC();
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
      NFB_EX_SYNC();
    }

    // ---------------- x0 = rgb_feat + direction_feat ----------------
    NFB_TC_WAIT();
#pragma unroll
    for (int c0 = 0; c0 < 48; c0 += 16) {
      float df[16];
      epi16(tl, c0, sf + F_B_DIR2 + c0, df);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c0 + j < NFB_ROW_CH) x[c0 + j] += df[j];
    }

    if (save) {
#pragma unroll
      for (int j = 0; j < 8; ++j) __stcs(sp + (SP_X0 + j) * GROUP, make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]));
      __stcs(sp + (SP_X0 + 8) * GROUP, make_float4(x[32], x[33], x[34], 0.f));
    }
    // ---------------- first pooling: weighted mean / variance over views ----------------
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(ex + tg * EXQ + 4 * j) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
    *reinterpret_cast<float4*>(ex + tg * EXQ + 32) = make_float4(x[32], x[33], x[34], w);
    NFB_EX_SYNC();
    if (active) {
      __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(mvps);
      __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(mvps + 36);
      pool4<9, POOL_MEAN_VAR>(ex + base * EXQ, V, v, 35, 1.f, [&](int q, const float4& m4, const float4& v4) {
        const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 4 * q + i;
          if (c < NFB_ROW_CH) {
            // the operand halves of this statistic, split once per sample instead of once per row
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
    NFB_EX_SYNC();

    // ---------------- base_fc.0 : [mean | var | x0] (105 -> 64) ----------------
    // operand words 0..34 = packed [mean | var] of the sample (shared memory), words 35..52 = x0 pairs, 53..55 = 0
    {
      // K chunks 0..3 (k < 64): statistics only
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        uint32_t hi[8], lo[8];
        const uint4 h0 = *reinterpret_cast<const uint4*>(mvps + 8 * kc), h1 = *reinterpret_cast<const uint4*>(mvps + 8 * kc + 4);
        hi[0] = h0.x; hi[1] = h0.y; hi[2] = h0.z; hi[3] = h0.w; hi[4] = h1.x; hi[5] = h1.y; hi[6] = h1.z; hi[7] = h1.w;
        if (NPASS == 3) {
          const uint4 l0 = *reinterpret_cast<const uint4*>(mvps + 36 + 8 * kc), l1 = *reinterpret_cast<const uint4*>(mvps + 36 + 8 * kc + 4);
          lo[0] = l0.x; lo[1] = l0.y; lo[2] = l0.z; lo[3] = l0.w; lo[4] = l1.x; lo[5] = l1.y; lo[6] = l1.z; lo[7] = l1.w;
        }
        a_store_words<NPASS>(tl, kc, hi, lo);
      }
      if (NPASS == 3) {               // A region holds K = 64 per round: issue the first round now
        NFB_TC_ISSUE(L_BASE0, 0, 4, false, 0);
        NFB_TC_WAIT();
      }
      constexpr int KC0 = (NPASS == 3) ? 4 : 0;   // chunk index offset of the second round
      {
        uint32_t hi[8], lo[8];
        hi[0] = mvps[32]; hi[1] = mvps[33]; hi[2] = mvps[34];
        if (NPASS == 3) { lo[0] = mvps[36 + 32]; lo[1] = mvps[36 + 33]; lo[2] = mvps[36 + 34]; }
#pragma unroll
        for (int j = 0; j < 5; ++j) pack_pair<NPASS>(x[2 * j], x[2 * j + 1], hi[3 + j], lo[3 + j]);
        a_store_words<NPASS>(tl, 4 - KC0, hi, lo);
#pragma unroll
        for (int j = 0; j < 8; ++j) pack_pair<NPASS>(x[10 + 2 * j], x[11 + 2 * j], hi[j], lo[j]);
        a_store_words<NPASS>(tl, 5 - KC0, hi, lo);
#pragma unroll
        for (int j = 0; j < 4; ++j) pack_pair<NPASS>(x[26 + 2 * j], x[27 + 2 * j], hi[j], lo[j]);
        pack_pair<NPASS>(x[34], 0.f, hi[4], lo[4]);
        hi[5] = hi[6] = hi[7] = 0u;
        lo[5] = lo[6] = lo[7] = 0u;
        a_store_words<NPASS>(tl, 6 - KC0, hi, lo);
      }
      if (NPASS == 3) NFB_TC_ISSUE(L_BASE0, 4, 7, true, 1);
      else NFB_TC_ISSUE(L_BASE0, 0, 7, false, 1);
      NFB_TC_WAIT();
    }

    // ---------------- base_fc.2 (64 -> 32) ----------------
#if NFB_VTC_PAIR
#pragma unroll
    for (int kc = 0; kc < 4; kc += 2) {
      float h[16], h2[16];
      uint32_t q[8], q2[8];
      epi16_pair<SAVE>(tl, 16 * kc, sf + F_B_BASE0 + 16 * kc, h, h2, q, q2);
      if (save) { st_codes8(sp, SP_H1 + 2 * kc, q); st_codes8(sp, SP_H1 + 2 * kc + 2, q2); }
      a_store16<NPASS>(tl, kc, h);
      a_store16<NPASS>(tl, kc + 1, h2);
    }
#else
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      float h[16];
      uint32_t q[8];
      epi16_s<SAVE>(tl, 16 * kc, sf + F_B_BASE0 + 16 * kc, h, q);
      if (save) st_codes8(sp, SP_H1 + 2 * kc, q);
      a_store16<NPASS>(tl, kc, h);
    }
#endif
    NFB_TC_ISSUE(L_BASE2, 0, 4, false, 2);
    NFB_TC_WAIT();

    // ---------------- vis_fc (32 -> 32 -> 33) on x1 * w ----------------
    float x1[32];
#if NFB_VTC_PAIR
    {
      float h[16], h2[16];
      uint32_t q[8], q2[8];
      epi16_pair<SAVE>(tl, 0, sf + F_B_BASE2, h, h2, q, q2);
      if (save) { st_codes8(sp, SP_X1, q); st_codes8(sp, SP_X1 + 2, q2); }
      float t[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { x1[j] = h[j]; t[j] = h[j] * w; }
      a_store16<NPASS>(tl, 0, t);
#pragma unroll
      for (int j = 0; j < 16; ++j) { x1[16 + j] = h2[j]; t[j] = h2[j] * w; }
      a_store16<NPASS>(tl, 1, t);
    }
#else
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      float h[16];
      uint32_t q[8];
      epi16_s<SAVE>(tl, 16 * kc, sf + F_B_BASE2 + 16 * kc, h, q);
      if (save) st_codes8(sp, SP_X1 + 2 * kc, q);
      float t[16];
      const float2 w2 = make_float2(w, w);
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        x1[16 * kc + j] = h[j]; x1[16 * kc + j + 1] = h[j + 1];
        const float2 r = __fmul2_rn(make_float2(h[j], h[j + 1]), w2);
        t[j] = r.x; t[j + 1] = r.y;
      }
      a_store16<NPASS>(tl, kc, t);
    }
#endif
    NFB_TC_ISSUE(L_VIS0, 0, 2, false, 1);
    NFB_TC_WAIT();
#if NFB_VTC_PAIR
    {
      float h[16], h2[16];
      uint32_t q[8], q2[8];
      epi16_pair<SAVE>(tl, 0, sf + F_B_VIS0, h, h2, q, q2);
      if (save) { st_codes8(sp, SP_HV, q); st_codes8(sp, SP_HV + 2, q2); }
      a_store16<NPASS>(tl, 0, h);
      a_store16<NPASS>(tl, 1, h2);
    }
#else
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      float h[16];
      uint32_t q[8];
      epi16_s<SAVE>(tl, 16 * kc, sf + F_B_VIS0 + 16 * kc, h, q);
      if (save) st_codes8(sp, SP_HV + 2 * kc, q);
      a_store16<NPASS>(tl, kc, h);
    }
#endif
    NFB_TC_ISSUE(L_VIS2, 0, 2, false, 2);
    NFB_TC_WAIT();

    // x2 = x1 + x_res ; vis1 = sigmoid(xv[32]) * mask ; vis_fc2 on x2 * vis1
    float vis1, sg1;
    uint32_t xvq16 = 0u;
    {
      float h[16];
      uint32_t q[8];
      epi16_s<SAVE>(tl, 32, sf + F_B_VIS2 + 32, h, q);
      if (SAVE) xvq16 = q[0];
      sg1 = sigmoid_f(h[0]);
      vis1 = sg1 * mk;
    }
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      float h[16];
      uint32_t q[8];
      epi16_s<SAVE>(tl, 16 * kc, sf + F_B_VIS2 + 16 * kc, h, q);
      if (save) st_codes8(sp, SP_XV + 2 * kc, q);
      float t[16];
      const float2 v2 = make_float2(vis1, vis1);
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float2 s2 = __fadd2_rn(make_float2(x1[16 * kc + j], x1[16 * kc + j + 1]), make_float2(h[j], h[j + 1]));      // x1 now holds x2
        x1[16 * kc + j] = s2.x; x1[16 * kc + j + 1] = s2.y;
        const float2 r = __fmul2_rn(s2, v2);
        t[j] = r.x; t[j + 1] = r.y;
      }
      if (save) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          __stcs(sp + (SP_X2 + 4 * kc + j) * GROUP, make_float4(x1[16 * kc + 4 * j], x1[16 * kc + 4 * j + 1], x1[16 * kc + 4 * j + 2], x1[16 * kc + 4 * j + 3]));
      }
      a_store16<NPASS>(tl, kc, t);
    }
    NFB_TC_ISSUE(L_VISB0, 0, 2, false, 3);
    NFB_TC_WAIT();
    float vis2, sg2;
    {
      float z = sf[F_B_VISB2];
      float2 za = make_float2(0.f, 0.f), zb = make_float2(0.f, 0.f);
#if NFB_VTC_PAIR
      {
        float h[16], h2[16];
        uint32_t q[8], q2[8];
        epi16_pair<SAVE>(tl, 0, sf + F_B_VISB0, h, h2, q, q2);
        if (save) { st_codes8(sp, SP_HV2, q); st_codes8(sp, SP_HV2 + 2, q2); }
#pragma unroll
        for (int j = 0; j < 16; ++j) z = fmaf(h[j], sf[F_W_VISB2 + j], z);
#pragma unroll
        for (int j = 0; j < 16; ++j) z = fmaf(h2[j], sf[F_W_VISB2 + 16 + j], z);
      }
#else
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        float h[16];
        uint32_t q[8];
        epi16_s<SAVE>(tl, 16 * kc, sf + F_B_VISB0 + 16 * kc, h, q);
        if (save) st_codes8(sp, SP_HV2 + 2 * kc, q);
        // 32-term dot product as four interleaved partial sums in two packed accumulators (a serial FFMA chain before)
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(sf + F_W_VISB2 + 16 * kc + j);
          za = __ffma2_rn(make_float2(h[j], h[j + 1]), make_float2(w4.x, w4.y), za);
          zb = __ffma2_rn(make_float2(h[j + 2], h[j + 3]), make_float2(w4.z, w4.w), zb);
        }
      }
      z += (za.x + za.y) + (zb.x + zb.y);
#endif
      sg2 = sigmoid_f(z);
      vis2 = sg2 * mk;
    }

    // ---------------- rgb_fc on [x2, vis2, ray_diff] (37 -> 16 -> 8 -> 1) ----------------
    {
      float t[16];
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
#pragma unroll
        for (int j = 0; j < 16; ++j) t[j] = x1[16 * kc + j];
        a_store16<NPASS>(tl, kc, t);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) t[j] = 0.f;
      t[0] = vis2; t[1] = rd[0]; t[2] = rd[1]; t[3] = rd[2]; t[4] = rd[3];
      a_store16<NPASS>(tl, 2, t);
    }
    NFB_TC_ISSUE(L_RGB0, 0, 3, false, 3);
    NFB_TC_WAIT();
    float logit;
    {
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
    }

    // ---------------- second pooling, blending, output ----------------
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(ex + tg * EXQ + 4 * j) = make_float4(x1[4 * j], x1[4 * j + 1], x1[4 * j + 2], x1[4 * j + 3]);
    *reinterpret_cast<float2*>(ex + tg * EXQ + 32) = make_float2(vis2, logit);
    *reinterpret_cast<float4*>(side + tg * SIDE) = make_float4(rgb_in0, rgb_in1, rgb_in2, 0.f);
    NFB_EX_SYNC();
    if (active) {
      float D = 1e-8f;
      for (int u = 0; u < V; ++u) D += ex[(base + u) * EXQ + 32];
      const float invD = 1.f / D;
      float* out = a.ps + (size_t)p * NFB_PS_STRIDE;
      pool4<8, POOL_MEAN_VAR>(ex + base * EXQ, V, v, 32, invD, [&](int q, const float4& m4, const float4& v4) {
        *reinterpret_cast<float4*>(out + PS_MEAN + 4 * q) = m4;
        *reinterpret_cast<float4*>(out + PS_VAR + 4 * q) = v4;
      });
      if (v == 0) {
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
    if (FUSED) NFB_EX_SYNC(); else named_bar_sync(bar_id, GROUP);            // exchange buffer is reused by the next tile
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*s_tmem, TMEM_ALLOC);
}

template <int NPASS, bool FUSED, bool SAVE = false>
int launch_view_tc_fwd(const ViewArgs& a, cudaStream_t st) {
  constexpr size_t smem = smem_bytes<NPASS>();
  cudaError_t e = cudaFuncSetAttribute(k_view_tc_fwd<NPASS, FUSED, SAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "k_view_tc_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int TS = row_map(a.V).TS;
  const int ntiles = (a.N + TS - 1) / TS;
  int grid = (ntiles + NG - 1) / NG;
  const int cap = nfb_num_sms();
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  k_view_tc_fwd<NPASS, FUSED, SAVE><<<grid, GROUP * NG, smem, st>>>(a);
  NFB_CHECK_LAUNCH("k_view_tc_fwd");
  return NFB_OK;
}

}  // namespace nfbvtc

// defined in nfb_view_tc_inst.cu (one instantiation per translation unit)
int nfb_launch_view_tc_fwd_p1_fused(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_tc_fwd_p3_fused(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_tc_fwd_p1_tensor(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_tc_fwd_p3_tensor(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_tc_fwd_p1_fused_save(const nfbview::ViewArgs& a, cudaStream_t st);
int nfb_launch_view_tc_fwd_p3_fused_save(const nfbview::ViewArgs& a, cudaStream_t st);
