// GNT dense layers on the 5th-generation tensor cores (tcgen05 + TMEM): one generic kernel for every 64-wide linear
// layer of the GNT network (gnt/transformer_network.py), in the operand scheme of nfb_view_tc.cuh:
//   * one thread per row (a (sample, view) row or a sample), 128-row groups = one M = 128 MMA tile,
//   * the thread writes its (optionally LayerNorm-ed) input row as bf16 hi / lo halves into TMEM (A operand),
//   * one thread of the group issues the K/16 MMAs against 64 x 64 weight tiles resident in shared memory
//     (canonical no-swizzle K-major layout) and commits to the group's mbarrier,
//   * every thread reads its fp32 accumulator row back for bias / ReLU / residual and stores it to HBM.
// NPASS = 3: x = hi + lo split of activations and weights, D = hi*hi + lo*hi + hi*lo in fp32 (fp32-equivalent);
// NPASS = 1: plain bf16 operands.
//
// Modes (what one launch computes per row; every weight is a 64 x 64 tile, torch layout [out][in]):
//   LIN_PRE   y0 = W0 LN(x)                                   view attention: qq = q_fc(attn_norm(q))
//   LIN_KV    k = W0 x ; v = W1 k ; pos = pos_fc(ray_diff) ;     view attention, per (sample, view) row: k = k_fc(F), v = v_fc(k);
//             y0 = v + pos ; y1[8] = ReLU(attn_fc.0(k - qq + pos))   k never leaves the SM: the row hands on v + pos and the 8 hidden
//                                                               units of attn_fc (288 B instead of 512 B per row)
//   LIN_QKV   y0 = W0 LN(x) ; y1 = W1 LN(x) ; y2 = W2 LN(x)   ray attention:  q, k, v projections
//   LIN_POST  y0 = W0 x + b + res                             out_fc + residual (view and ray attention)
//   LIN_FFN   y0 = W2 ReLU(W1 LN(x) + b1) + b2 + x            feed-forward block: fc1 [256][64] = 4 N-chunks, fc2 [64][256] = 4 K-chunks
//   LIN_QFC   y0 = W1 ReLU(W0 [x | posenc(pts) | posenc(dir)] + b0) + b1   q_fc of the even layers: fc.0 [64][190] = 3 K-rounds
//   LIN_EMBED y0 = W1 ReLU(W0 x35 + b0) + b1                   rgbfeat_fc on the 35-channel rows (W0 [64][35] zero-padded to K = 64)
#pragma once
#include "nfb_common.cuh"
#include "nfb_tc.cuh"

namespace gnttc {
using namespace nfbtc;

constexpr int GROUP = 128;
constexpr int TD = 64;                           // tile edge = netwidth
constexpr int TILE_BYTES = TD * TD * 2;          // one bf16 64 x 64 tile
enum : int { LIN_PRE = 0, LIN_KV = 1, LIN_QKV = 2, LIN_POST = 3, LIN_FFN = 4, LIN_EMBED = 5, LIN_QFC = 6 };

__host__ __device__ constexpr int mode_tiles(int mode) { return mode == LIN_PRE || mode == LIN_POST ? 1 : (mode == LIN_KV || mode == LIN_EMBED) ? 2 : mode == LIN_QKV ? 3 : mode == LIN_QFC ? 4 : 8; }
__host__ __device__ constexpr int mode_groups(int mode) { return mode == LIN_FFN ? 2 : 4; }       // FFN needs 256 TMEM columns per group
__host__ __device__ constexpr int mode_cols(int mode) { return mode == LIN_FFN ? 256 : 128; }
// TMEM columns of a group: D0 [0,64) | A hi [64,96) | A lo [96,128) | FFN only: D1 [128,192) | A2 hi [192,224) | A2 lo [224,256)
constexpr int C_D0 = 0, C_A = 64, C_ALO = 96, C_D1 = 128, C_A2 = 192, C_A2LO = 224;

// fp32 side tables of LIN_KV (pos_fc and attn_fc.0 stay on the CUDA cores: 4 -> 8 -> 64 and 64 -> 8)
enum : int { KV_P0 = 0 /*[4][8]*/, KV_P0_B = 32, KV_P2 = 40 /*[8][64]*/, KV_P2_B = KV_P2 + 8 * TD, KV_A0 = KV_P2_B + TD /*[64][8]*/,
             KV_A0_B = KV_A0 + TD * 8, KV_TOTAL = KV_A0_B + 8 };

struct LinArgs {
  long long M;                 // rows
  const float* x;              // [M][64] input rows
  const float* res;            // LIN_POST: residual rows [M][64]
  float* y0; float* y1; float* y2;
  const float* w[3];           // 64 x 64 tiles (LIN_FFN: w[0] = fc1 [256][64], w[1] = fc2 [64][256])
  const float* b0;             // LIN_POST: bias [64]; LIN_FFN: fc1 bias [256]
  const float* b1;             // LIN_FFN: fc2 bias [64]
  const float* ln_w; const float* ln_b;   // LayerNorm of the input (LIN_PRE, LIN_QKV, LIN_FFN), eps 1e-6
  const float* pts; const float* ray_d; int S;   // LIN_QFC: sample positions [M][3], ray directions [M / S][3]
  // LIN_KV: qq [M / V][64], ray_diff [M][4], pos_fc.0 (w [8][4], b), pos_fc.2 (w [64][8], b), attn_fc.0 (w [8][64], b)
  const float* qq; const float* ray_diff; int V;
  const float* p0_w; const float* p0_b; const float* p2_w; const float* p2_b; const float* a0_w; const float* a0_b;
};

// Embedder (transformer_network.py:6-37): [x, sin(x f0), cos(x f0), sin(x f1), ...], f_k = 2^k, k = 0..9 -> 63 values (+ 1 pad)
__device__ __forceinline__ void posenc64(const float (&x3)[3], float (&e)[TD]) {
  e[0] = x3[0]; e[1] = x3[1]; e[2] = x3[2];
#pragma unroll
  for (int f = 0; f < 10; ++f) {
    const float fr = (float)(1 << f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float ang = __fmul_rn(x3[i], fr);
      e[3 + 6 * f + i] = sinf(ang);
      e[3 + 6 * f + 3 + i] = cosf(ang);
    }
  }
  e[63] = 0.f;
}

// Staging rows of the warp-coalesced row I/O (coop_row_load / coop_row_store): row stride 272 B keeps the 16-byte row reads of a
// warp conflict-free.  (Round 1 used them to prefetch LIN_KV's next tile with cp.async, which measured +1 %; coalescing the loads
// and the 256-byte v + pos stores through them is worth more.)
constexpr int ROW_STAGE_STRIDE = 272;
// every mode stages its 64-float rows through shared memory so that global loads / stores are warp-coalesced (see coop_row_load)
__host__ __device__ constexpr size_t stage_bytes(int mode) { return (size_t)mode_groups(mode) * GROUP * ROW_STAGE_STRIDE; }

template <int NPASS, int MODE>
__host__ __device__ constexpr size_t lin_smem_bytes() {
  return (size_t)mode_tiles(MODE) * TILE_BYTES * (NPASS == 3 ? 2 : 1) + sizeof(float) * (256 + 64 + 128 + KV_TOTAL) + stage_bytes(MODE) +
         mode_groups(MODE) * 8 + 16;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__host__ __device__ constexpr uint32_t canon_off(int n, int k) {      // element (n, k) of a [64][64] K-major tile
  return (uint32_t)((k >> 3) * (TD * 16) + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2);
}

// w: element (n, k) at w[n * ldw + k]
template <int NPASS>
static __device__ void load_tile64(uint8_t* hi, uint8_t* lo, const float* __restrict__ w, int ldw, int tid, int nt, int k_real = TD) {
  for (int i = tid; i < TD * TD; i += nt) {
    const int n = i >> 6, k = i & 63;
    const float v = k < k_real ? __ldg(w + (size_t)n * ldw + k) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const uint32_t off = canon_off(n, k);
    *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
    if (NPASS == 3) *reinterpret_cast<__nv_bfloat16*>(lo + off) = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// this thread's 64-value row -> A operand (4 K-chunks of 16) at column base ca (hi) / calo (lo)
template <int NPASS>
__device__ __forceinline__ void a_store_row(uint32_t tl, int ca, int calo, const float (&v)[TD]) {
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (NPASS == 3) split_bf16(v[16 * kc + 2 * j], v[16 * kc + 2 * j + 1], hi[j], lo[j]);
      else hi[j] = pack_bf16(v[16 * kc + 2 * j], v[16 * kc + 2 * j + 1]);
    }
    tmem_st8(tl + ca + 8 * kc, hi);
    if (NPASS == 3) tmem_st8(tl + calo + 8 * kc, lo);
  }
}

// D[dcol .. dcol+64) (+)= A[64 values at ca / calo] x tile^T ; one thread
template <int NPASS>
__device__ __forceinline__ void issue_tile(uint32_t tb, int dcol, int ca, int calo, uint32_t tile_hi, uint32_t tile_lo, bool acc0) {
  constexpr uint32_t idesc = idesc_bf16(128, TD);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint64_t dh = smem_desc(tile_hi + ks * 2 * TD * 16, TD * 16, 128);
    mma_ts(tb + dcol, tb + ca + 8 * ks, dh, idesc, acc0 || ks > 0);
    if (NPASS == 3) {
      const uint64_t dl = smem_desc(tile_lo + ks * 2 * TD * 16, TD * 16, 128);
      mma_ts(tb + dcol, tb + calo + 8 * ks, dh, idesc, true);
      mma_ts(tb + dcol, tb + ca + 8 * ks, dl, idesc, true);
    }
  }
}

__device__ __forceinline__ void d_load_row(uint32_t tl, int dcol, float (&y)[TD]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float t[16];
    tmem_ld16(tl + dcol + 16 * c, t);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) y[16 * c + j] = t[j];
  }
}

__device__ __forceinline__ void row_load(const float* __restrict__ p, float (&x)[TD]) {
#pragma unroll
  for (int c = 0; c < TD; c += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + c));
    x[c] = t.x; x[c + 1] = t.y; x[c + 2] = t.z; x[c + 3] = t.w;
  }
}
__device__ __forceinline__ void row_store(float* __restrict__ p, const float (&x)[TD]) {
#pragma unroll
  for (int c = 0; c < TD; c += 4) *reinterpret_cast<float4*>(p + c) = make_float4(x[c], x[c + 1], x[c + 2], x[c + 3]);
}
__device__ __forceinline__ void row_layer_norm(float (&x)[TD], const float* __restrict__ w, const float* __restrict__ b) {
  float mu = 0.f;
#pragma unroll
  for (int c = 0; c < TD; ++c) mu += x[c];
  mu *= (1.f / TD);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < TD; ++c) var = fmaf(x[c] - mu, x[c] - mu, var);
  var *= (1.f / TD);
  const float rstd = 1.f / sqrtf(var + 1e-6f);
#pragma unroll
  for (int c = 0; c < TD; ++c) x[c] = fmaf((x[c] - mu) * rstd, w[c], b[c]);
}

#define GNT_TC_ISSUE(DCOL, CA, CALO, TILE, ACC0)                                                        \
  do {                                                                                                  \
    tmem_st_wait();                                                                                     \
    fence_before_sync();                                                                                \
    named_bar_sync(bar_id, GROUP);                                                                      \
    if (tg == 0) {                                                                                      \
      fence_after_sync();                                                                               \
      issue_tile<NPASS>(tb, DCOL, CA, CALO, sB_addr + (TILE) * TILE_BYTES, sB_addr + (NT + (TILE)) * TILE_BYTES, ACC0); \
      mma_commit(mbar);                                                                                 \
    }                                                                                                   \
  } while (0)
#define GNT_TC_WAIT()          \
  do {                         \
    mbar_wait(mbar, phase);    \
    phase ^= 1u;               \
    fence_after_sync();        \
  } while (0)

template <int NPASS, int MODE>
__global__ void __launch_bounds__(GROUP * mode_groups(MODE), 1) k_gnt_lin_tc(LinArgs a) {
  constexpr int NG = mode_groups(MODE), GC = mode_cols(MODE), NT = mode_tiles(MODE);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sB = smem_raw;                                                    // NT hi tiles, then NT lo tiles
  float* sf = reinterpret_cast<float*>(smem_raw + (size_t)NT * TILE_BYTES * (NPASS == 3 ? 2 : 1));
  float* s_b0 = sf;            // 256
  float* s_b1 = sf + 256;      // 64
  float* s_ln = sf + 320;      // 64 + 64
  float* s_kv = sf + 448;      // KV_TOTAL
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(sf + 448 + KV_TOTAL);          // LIN_KV: NG x [128 rows][272 B]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + stage_bytes(MODE));
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + NG);

  const int tid = threadIdx.x, warp = tid >> 5, nt = blockDim.x;
  const int grp = tid / GROUP, tg = tid % GROUP;
  const int bar_id = 1 + grp;
  uint64_t* mbar = s_bar + grp;

  if (warp == 0) tmem_alloc(s_tmem, NG * GC);
  if (tid == 0) {
    for (int g = 0; g < NG; ++g) mbar_init(s_bar + g, 1);
    mbar_init_fence();
  }
  if (MODE == LIN_FFN) {
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      load_tile64<NPASS>(sB + j * TILE_BYTES, sB + (NT + j) * TILE_BYTES, a.w[0] + (size_t)j * TD * TD, TD, tid, nt);        // fc1 rows 64j..
      load_tile64<NPASS>(sB + (4 + j) * TILE_BYTES, sB + (NT + 4 + j) * TILE_BYTES, a.w[1] + j * TD, 4 * TD, tid, nt);         // fc2 cols 64j..
    }
    for (int i = tid; i < 256; i += nt) s_b0[i] = __ldg(a.b0 + i);
    for (int i = tid; i < 64; i += nt) s_b1[i] = __ldg(a.b1 + i);
  } else if (MODE == LIN_QFC) {
    // fc.0 [64][190]: columns [0,64) = q, [64,127) = posenc(pts), [127,190) = posenc(dir); the 63-wide blocks are zero-padded
    load_tile64<NPASS>(sB, sB + NT * TILE_BYTES, a.w[0], 190, tid, nt);
    load_tile64<NPASS>(sB + TILE_BYTES, sB + (NT + 1) * TILE_BYTES, a.w[0] + 64, 190, tid, nt, 63);
    load_tile64<NPASS>(sB + 2 * TILE_BYTES, sB + (NT + 2) * TILE_BYTES, a.w[0] + 127, 190, tid, nt, 63);
    load_tile64<NPASS>(sB + 3 * TILE_BYTES, sB + (NT + 3) * TILE_BYTES, a.w[1], TD, tid, nt);
    for (int i = tid; i < 64; i += nt) { s_b0[i] = __ldg(a.b0 + i); s_b1[i] = __ldg(a.b1 + i); }
  } else if (MODE == LIN_EMBED) {
    load_tile64<NPASS>(sB, sB + NT * TILE_BYTES, a.w[0], NFB_ROW_CH, tid, nt, NFB_ROW_CH);
    load_tile64<NPASS>(sB + TILE_BYTES, sB + (NT + 1) * TILE_BYTES, a.w[1], TD, tid, nt);
    for (int i = tid; i < 64; i += nt) { s_b0[i] = __ldg(a.b0 + i); s_b1[i] = __ldg(a.b1 + i); }
  } else {
#pragma unroll 1
    for (int j = 0; j < NT; ++j) load_tile64<NPASS>(sB + j * TILE_BYTES, sB + (NT + j) * TILE_BYTES, a.w[j], TD, tid, nt);
    if (MODE == LIN_POST)
      for (int i = tid; i < 64; i += nt) s_b0[i] = __ldg(a.b0 + i);
  }
  if (MODE == LIN_KV) {
    for (int i = tid; i < 32; i += nt) s_kv[KV_P0 + i] = __ldg(a.p0_w + (i & 7) * 4 + (i >> 3));          // [k][j] <- w[j][k]
    for (int i = tid; i < 8; i += nt) { s_kv[KV_P0_B + i] = __ldg(a.p0_b + i); s_kv[KV_A0_B + i] = __ldg(a.a0_b + i); }
    for (int i = tid; i < 8 * TD; i += nt) {
      s_kv[KV_P2 + i] = __ldg(a.p2_w + (i & 63) * 8 + (i >> 6));                                             // [j][c] <- w[c][j]
      s_kv[KV_A0 + i] = __ldg(a.a0_w + (i & 7) * TD + (i >> 3));                                             // [c][j] <- w[j][c]
    }
    for (int i = tid; i < TD; i += nt) s_kv[KV_P2_B + i] = __ldg(a.p2_b + i);
  }
  if (MODE == LIN_PRE || MODE == LIN_QKV || MODE == LIN_FFN)
    for (int i = tid; i < 64; i += nt) { s_ln[i] = __ldg(a.ln_w + i); s_ln[64 + i] = __ldg(a.ln_b + i); }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const uint32_t tb = *s_tmem + (uint32_t)(grp * GC);
  const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t sB_addr = smem_u32(sB);
  uint32_t phase = 0;
  const long long ntiles = (a.M + GROUP - 1) / GROUP;

  uint8_t* my_stage = s_stage + ((size_t)grp * GROUP + tg) * ROW_STAGE_STRIDE;
  const int lane = tid & 31;
  uint8_t* warp_stage = my_stage - (size_t)lane * ROW_STAGE_STRIDE;
  // Row I/O.  A thread's own 256-byte row is 16 LDG.128 / STG.128 that each touch 32 different lines across the warp
  // (the uncoalesced pattern of the IBRNet gather); instead the warp moves its 32 contiguous rows as 16 fully coalesced 512-byte
  // requests and the rows are transposed through the staging buffer.
  auto coop_row_load = [&](const float* __restrict__ base, long long row0, bool act, float (&x)[TD]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int piece = 32 * i + lane, r = piece >> 4, c = piece & 15;
      if (row0 + r < a.M) cp_async16(warp_stage + (size_t)r * ROW_STAGE_STRIDE + 16 * c, base + (row0 + r) * TD + 4 * c);
    }
    cp_async_wait_all();
    __syncwarp();
#pragma unroll
    for (int c = 0; c < TD; c += 4) {
      const float4 t4 = act ? *reinterpret_cast<const float4*>(my_stage + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      x[c] = t4.x; x[c + 1] = t4.y; x[c + 2] = t4.z; x[c + 3] = t4.w;
    }
    __syncwarp();
  };
  auto coop_row_store = [&](float* __restrict__ base, long long row0, const float (&y)[TD]) {
#pragma unroll
    for (int c = 0; c < TD; c += 4) *reinterpret_cast<float4*>(my_stage + 4 * c) = make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int piece = 32 * i + lane, r = piece >> 4, c = piece & 15;
      if (row0 + r < a.M)
        *reinterpret_cast<float4*>(base + (row0 + r) * TD + 4 * c) = *reinterpret_cast<const float4*>(warp_stage + (size_t)r * ROW_STAGE_STRIDE + 16 * c);
    }
    __syncwarp();
  };
  for (long long tile = (long long)blockIdx.x * NG + grp; tile < ntiles; tile += (long long)gridDim.x * NG) {
    const long long row = tile * GROUP + tg;
    const bool active = row < a.M;
    float x[TD];
    if (MODE == LIN_EMBED) {
#pragma unroll
      for (int c = 0; c < TD; ++c) x[c] = (active && c < NFB_ROW_CH) ? __ldg(a.x + row * NFB_ROW_CH + c) : 0.f;
    } else coop_row_load(a.x, row - lane, active, x);
    if (MODE == LIN_PRE || MODE == LIN_QKV) row_layer_norm(x, s_ln, s_ln + 64);
    if (MODE == LIN_FFN) {
      float xn[TD];
#pragma unroll
      for (int c = 0; c < TD; ++c) xn[c] = x[c];
      row_layer_norm(xn, s_ln, s_ln + 64);
      a_store_row<NPASS>(tl, C_A, C_ALO, xn);
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        GNT_TC_ISSUE(C_D0, C_A, C_ALO, j, false);                 // h_j = fc1[64j .. 64j+64) . LN(x)
        GNT_TC_WAIT();
        float h[TD];
        d_load_row(tl, C_D0, h);
#pragma unroll
        for (int c = 0; c < TD; ++c) h[c] = fmaxf(h[c] + s_b0[64 * j + c], 0.f);
        a_store_row<NPASS>(tl, C_A2, C_A2LO, h);
        GNT_TC_ISSUE(C_D1, C_A2, C_A2LO, 4 + j, j > 0);           // y += fc2[:, 64j .. 64j+64) . h_j
        GNT_TC_WAIT();
      }
      float y[TD];
      d_load_row(tl, C_D1, y);
#pragma unroll
      for (int c = 0; c < TD; ++c) y[c] += s_b1[c] + x[c];
      coop_row_store(a.y0, row - lane, y);
    } else {
      a_store_row<NPASS>(tl, C_A, C_ALO, x);
      GNT_TC_ISSUE(C_D0, C_A, C_ALO, 0, false);
      GNT_TC_WAIT();
      float y[TD];
      d_load_row(tl, C_D0, y);
      if (MODE == LIN_POST) {
        float r[TD];
        coop_row_load(a.res, row - lane, active, r);
#pragma unroll
        for (int c = 0; c < TD; ++c) y[c] += s_b0[c] + r[c];
      }
      if (MODE == LIN_QFC) {
        // two more K-rounds into the same accumulator: the positional encodings of the point and of the view direction
        const long long rs = active ? row : 0;
        {
          const float p3[3] = {__ldg(a.pts + rs * 3), __ldg(a.pts + rs * 3 + 1), __ldg(a.pts + rs * 3 + 2)};
          float e[TD];
          posenc64(p3, e);
          a_store_row<NPASS>(tl, C_A, C_ALO, e);
          GNT_TC_ISSUE(C_D0, C_A, C_ALO, 1, true);
          GNT_TC_WAIT();
        }
        {
          const long long r = rs / a.S;
          const float dx = __ldg(a.ray_d + r * 3), dy = __ldg(a.ray_d + r * 3 + 1), dz = __ldg(a.ray_d + r * 3 + 2);
          const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
          const float d3[3] = {__fdiv_rn(dx, nrm), __fdiv_rn(dy, nrm), __fdiv_rn(dz, nrm)};
          float e[TD];
          posenc64(d3, e);
          a_store_row<NPASS>(tl, C_A, C_ALO, e);
          GNT_TC_ISSUE(C_D0, C_A, C_ALO, 2, true);
          GNT_TC_WAIT();
        }
        d_load_row(tl, C_D0, y);
#pragma unroll
        for (int c = 0; c < TD; ++c) y[c] = fmaxf(y[c] + s_b0[c], 0.f);
        a_store_row<NPASS>(tl, C_A, C_ALO, y);
        GNT_TC_ISSUE(C_D0, C_A, C_ALO, 3, false);
        GNT_TC_WAIT();
        d_load_row(tl, C_D0, y);
#pragma unroll
        for (int c = 0; c < TD; ++c) y[c] += s_b1[c];
      }
      if (MODE == LIN_EMBED) {
#pragma unroll
        for (int c = 0; c < TD; ++c) y[c] = fmaxf(y[c] + s_b0[c], 0.f);
        a_store_row<NPASS>(tl, C_A, C_ALO, y);
        GNT_TC_ISSUE(C_D0, C_A, C_ALO, 1, false);
        GNT_TC_WAIT();
        d_load_row(tl, C_D0, y);
#pragma unroll
        for (int c = 0; c < TD; ++c) y[c] += s_b1[c];
      }
      if (MODE != LIN_KV) coop_row_store(a.y0, row - lane, y);
      if (MODE == LIN_KV) {
        a_store_row<NPASS>(tl, C_A, C_ALO, y);                     // v = v_fc(k): the projected k is the next input
        GNT_TC_ISSUE(C_D0, C_A, C_ALO, 1, false);
        // while the MMA runs: pos = pos_fc(ray_diff) and the hidden units of attn_fc on k - qq + pos (CUDA cores)
        const long long rs = active ? row : 0;
        float pos[TD];
        {
          const float4 rd4 = __ldg(reinterpret_cast<const float4*>(a.ray_diff) + rs);
          float p8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            p8[j] = fmaxf(s_kv[KV_P0_B + j] + rd4.x * s_kv[KV_P0 + j] + rd4.y * s_kv[KV_P0 + 8 + j] + rd4.z * s_kv[KV_P0 + 16 + j] +
                          rd4.w * s_kv[KV_P0 + 24 + j], 0.f);
#pragma unroll
          for (int c = 0; c < TD; ++c) pos[c] = s_kv[KV_P2_B + c];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int c = 0; c < TD; c += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(s_kv + KV_P2 + j * TD + c);
              pos[c] = fmaf(p8[j], w4.x, pos[c]); pos[c + 1] = fmaf(p8[j], w4.y, pos[c + 1]);
              pos[c + 2] = fmaf(p8[j], w4.z, pos[c + 2]); pos[c + 3] = fmaf(p8[j], w4.w, pos[c + 3]);
            }
          }
        }
        float a8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a8[j] = s_kv[KV_A0_B + j];
        {
          const float4* qr = reinterpret_cast<const float4*>(a.qq + (rs / a.V) * TD);
#pragma unroll
          for (int c = 0; c < TD; c += 4) {
            const float4 q4 = __ldg(qr + c / 4);
            const float t0 = y[c] - q4.x + pos[c], t1 = y[c + 1] - q4.y + pos[c + 1], t2 = y[c + 2] - q4.z + pos[c + 2],
                        t3 = y[c + 3] - q4.w + pos[c + 3];
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
              const float4 w0 = *reinterpret_cast<const float4*>(s_kv + KV_A0 + c * 8 + j);
              const float4 w1 = *reinterpret_cast<const float4*>(s_kv + KV_A0 + (c + 1) * 8 + j);
              const float4 w2 = *reinterpret_cast<const float4*>(s_kv + KV_A0 + (c + 2) * 8 + j);
              const float4 w3 = *reinterpret_cast<const float4*>(s_kv + KV_A0 + (c + 3) * 8 + j);
              a8[j] += t0 * w0.x + t1 * w1.x + t2 * w2.x + t3 * w3.x;
              a8[j + 1] += t0 * w0.y + t1 * w1.y + t2 * w2.y + t3 * w3.y;
              a8[j + 2] += t0 * w0.z + t1 * w1.z + t2 * w2.z + t3 * w3.z;
              a8[j + 3] += t0 * w0.w + t1 * w1.w + t2 * w2.w + t3 * w3.w;
            }
          }
        }
        GNT_TC_WAIT();
        d_load_row(tl, C_D0, y);
#pragma unroll
        for (int c = 0; c < TD; ++c) y[c] += pos[c];
        coop_row_store(a.y0, row - lane, y);          // v + pos: 256 B per row, coalesced through the staging rows
        if (active) {
          float4* o8 = reinterpret_cast<float4*>(a.y1 + row * 8);
          o8[0] = make_float4(fmaxf(a8[0], 0.f), fmaxf(a8[1], 0.f), fmaxf(a8[2], 0.f), fmaxf(a8[3], 0.f));
          o8[1] = make_float4(fmaxf(a8[4], 0.f), fmaxf(a8[5], 0.f), fmaxf(a8[6], 0.f), fmaxf(a8[7], 0.f));
        }
      }
      if (MODE == LIN_QKV) {
        GNT_TC_ISSUE(C_D0, C_A, C_ALO, 1, false);                  // A (LN(x)) is unchanged
        GNT_TC_WAIT();
        d_load_row(tl, C_D0, y);
        coop_row_store(a.y1, row - lane, y);
        GNT_TC_ISSUE(C_D0, C_A, C_ALO, 2, false);
        GNT_TC_WAIT();
        d_load_row(tl, C_D0, y);
        coop_row_store(a.y2, row - lane, y);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*s_tmem, NG * GC);
}

template <int NPASS, int MODE>
int launch_lin(const LinArgs& a, cudaStream_t st, const char* name) {
  constexpr size_t smem = lin_smem_bytes<NPASS, MODE>();
  cudaError_t e = cudaFuncSetAttribute(k_gnt_lin_tc<NPASS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
  constexpr int NG = mode_groups(MODE);
  const long long ntiles = (a.M + GROUP - 1) / GROUP;
  long long grid = (ntiles + NG - 1) / NG;
  const int cap = nfb_num_sms();
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  k_gnt_lin_tc<NPASS, MODE><<<(int)grid, GROUP * NG, smem, st>>>(a);
  NFB_CHECK_LAUNCH(name);
  return NFB_OK;
}

}  // namespace gnttc
