// Stand-alone probe of the tcgen05 building blocks in nfb_tc.cuh (run on a B200 through gpurun):
//   128 x N x K bf16 GEMM tiles, A from TMEM (.ts) or shared memory (.ss), B from shared memory in the canonical
//   no-swizzle K-major layout, D read back with tcgen05.ld, compared with a host reference.
// Usage: probe_tcgen05 <variant>   (one variant per process so that a trap in one does not poison the others)
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o probe_tcgen05 tests/probes/probe_tcgen05.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../nerfool_b200/csrc/nfb_tc.cuh"

using namespace nfbtc;

struct Cfg {
  int N, K;        // K multiple of 16
  int a_in_tmem;   // 1: .ts   0: .ss
  int swap_lbo;    // 1: row-group-major storage (LBO=128)   2: K-chunk-major storage, LBO/SBO fields exchanged (diagnostic)
  int swap_pack;   // 1: (high, low) packing of A in TMEM (diagnostic)
  int passes;      // 1: hi only   3: hi*hi + lo*hi + hi*lo
  int ones_bias;   // 1: bias through a persistent "ones" K-block
  int groups;      // row groups (128 threads each) running concurrently in the CTA
  int b_mn;        // 0: B stored K-major.  1/2: B' = W^T read MN-major from the K-major tile of W[K][N] (1: LBO=128,SBO=K*16  2: exchanged)
};

// canonical K-major no-swizzle offset (bytes) of element (r, k) of an [R][K] bf16 tile stored K-chunk-major
__host__ __device__ inline uint32_t canon_off(int r, int k, int R) {
  return (uint32_t)((k >> 3) * (R * 16) + (r >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

__global__ void __launch_bounds__(512, 1)
k_probe(Cfg c, const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ bias,
        float* __restrict__ D, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar[4];
  const int tid = threadIdx.x, grp = tid / 128, tg = tid % 128, warp = tid / 32;
  const int N = c.N, K = c.K;
  const int KB = K + 16;                              // + ones block
  // smem carve-up: B_hi [N][KB], B_lo [N][KB], then per group A_hi [128][K], A_lo [128][K] (ss mode)
  uint8_t* sBhi = smem;
  uint8_t* sBlo = sBhi + N * KB * 2;
  uint8_t* sA = sBlo + N * KB * 2;
  const int a_tile_bytes = 128 * K * 2;
  uint8_t* sAhi = sA + grp * 2 * a_tile_bytes;
  uint8_t* sAlo = sAhi + a_tile_bytes;

  if (warp == 0) tmem_alloc(&s_tmem, 512);
  if (tid == 0) {
    for (int g = 0; g < 4; ++g) mbar_init(&s_bar[g], 1);
    mbar_init_fence();
  }
  // weights -> canonical layout (hi / lo split), bias into column K of the extra block
  for (int i = tid; i < N * KB; i += blockDim.x) {
    const int n = i / KB, k = i % KB;
    float w = 0.f;
    if (k < K) w = B[n * K + k];
    else if (k == K && c.ones_bias) w = bias[n];
    const __nv_bfloat16 h = __float2bfloat16_rn(w);
    const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
    uint32_t off = c.swap_lbo == 1 ? (uint32_t)((k >> 3) * 128 + (n >> 3) * (KB * 16) + (n & 7) * 16 + (k & 7) * 2)
                              : canon_off(n, k, N);
    if (c.b_mn) off = canon_off(k, n, KB);          // physical tile = W[k][n] (rows k, KB of them), K-major canonical
    *reinterpret_cast<__nv_bfloat16*>(sBhi + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(sBlo + off) = l;
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = s_tmem;
  if (grp < c.groups) {
    // per-group TMEM columns: D [0,64) | A_hi [64, 64+K/2) | A_lo [.., +K/2) (3-pass only) | ones (8 cols)
    const uint32_t gcol = (uint32_t)grp * (uint32_t)(512 / c.groups);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t d_col = tbase + gcol;
    const uint32_t ahi_col = d_col + (N > 64 ? 128 : 64), alo_col = ahi_col + K / 2;
    const uint32_t ones_col = (c.passes == 3) ? alo_col + K / 2 : alo_col;

    // ---- A operand: this thread's row ----
    const float* arow = A + ((size_t)grp * 128 + tg) * K;
    if (c.a_in_tmem) {
      for (int k0 = 0; k0 < K; k0 += 16) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = arow[k0 + 2 * j], b = arow[k0 + 2 * j + 1];
          if (c.swap_pack) { float t = a; a = b; b = t; }
          split_bf16(a, b, hi[j], lo[j]);
        }
        tmem_st8(lane_base + ahi_col + k0 / 2, hi);
        if (c.passes == 3) tmem_st8(lane_base + alo_col + k0 / 2, lo);
      }
      uint32_t ones[8] = {0x00003F80u, 0, 0, 0, 0, 0, 0, 0};   // bf16(1.0) in element 0
      if (c.swap_pack) ones[0] = 0x3F800000u;
      tmem_st8(lane_base + ones_col, ones);
      tmem_st_wait();
    } else {
      for (int k = 0; k < K; ++k) {
        const float a = arow[k];
        const __nv_bfloat16 h = __float2bfloat16_rn(a);
        const __nv_bfloat16 l = __float2bfloat16_rn(a - __bfloat162float(h));
        uint32_t off = c.swap_lbo == 1 ? (uint32_t)((k >> 3) * 128 + (tg >> 3) * (K * 16) + (tg & 7) * 16 + (k & 7) * 2)
                                  : canon_off(tg, k, 128);
        *reinterpret_cast<__nv_bfloat16*>(sAhi + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(sAlo + off) = l;
      }
      fence_proxy_async_smem();
    }
    fence_before_sync();
    asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");

    if (tg == 0) {
      fence_after_sync();
      const uint32_t idesc = idesc_bf16(128, N) | (c.b_mn ? (1u << 16) : 0u);
      uint32_t b_lbo = c.swap_lbo == 1 ? 128u : (uint32_t)N * 16u, b_sbo = c.swap_lbo == 1 ? (uint32_t)KB * 16u : 128u;
      uint32_t a_lbo = c.swap_lbo == 1 ? 128u : 128u * 16u, a_sbo = c.swap_lbo == 1 ? (uint32_t)K * 16u : 128u;
      const uint32_t b_kstep = c.swap_lbo == 1 ? 256u : 2u * (uint32_t)N * 16u;   // bytes per K=16
      const uint32_t a_kstep = c.swap_lbo == 1 ? 256u : 2u * 128u * 16u;
      if (c.swap_lbo == 2) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
      uint32_t b_kstep_eff = b_kstep;
      if (c.b_mn) {
        b_lbo = 128u; b_sbo = (uint32_t)KB * 16u; b_kstep_eff = 256u;
        if (c.b_mn == 2) { b_lbo = (uint32_t)KB * 16u; b_sbo = 128u; }
      }
      bool acc = false;
      for (int pass = 0; pass < c.passes; ++pass) {
        const uint8_t* bsel = (pass == 2) ? sBlo : sBhi;        // pass 0: hi*hi, 1: lo*hi, 2: hi*lo
        for (int ks = 0; ks < K / 16; ++ks) {
          const uint64_t bd = smem_desc(smem_u32(bsel) + ks * b_kstep_eff, b_lbo, b_sbo);
          if (c.a_in_tmem) {
            const uint32_t acol = ((pass == 1) ? alo_col : ahi_col) + ks * 8;
            mma_ts(d_col, acol, bd, idesc, acc);
          } else {
            const uint8_t* asel = (pass == 1) ? sAlo : sAhi;
            const uint64_t ad = smem_desc(smem_u32(asel) + ks * a_kstep, a_lbo, a_sbo);
            mma_ss(d_col, ad, bd, idesc, acc);
          }
          acc = true;
        }
      }
      if (c.ones_bias && c.a_in_tmem) {
        const int ks = K / 16;
        mma_ts(d_col, ones_col, smem_desc(smem_u32(sBhi) + ks * b_kstep_eff, b_lbo, b_sbo), idesc, true);
        if (c.passes == 3) mma_ts(d_col, ones_col, smem_desc(smem_u32(sBlo) + ks * b_kstep_eff, b_lbo, b_sbo), idesc, true);
      }
      mma_commit(&s_bar[grp]);
    }
    // bounded wait
    int spins = 0;
    while (!mbar_try_wait(&s_bar[grp], 0)) {
      if (++spins > (1 << 22)) { if (tg == 0) atomicExch(status, 1); break; }
    }
    fence_after_sync();
    for (int n0 = 0; n0 < N; n0 += 16) {
      float v[16];
      tmem_ld16(lane_base + d_col + n0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) D[((size_t)grp * 128 + tg) * N + n0 + j] = v[j];
    }
    fence_before_sync();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  // N, K, a_in_tmem, swap_lbo, swap_pack, passes, ones_bias, groups
  const Cfg table[] = {
      {48, 32, 1, 0, 0, 1, 0, 1, 0},    // 0: ts, canonical
      {48, 32, 1, 1, 0, 1, 0, 1, 0},    // 1: ts, LBO/SBO swapped
      {48, 32, 1, 0, 1, 1, 0, 1, 0},    // 2: ts, A packing swapped
      {48, 32, 0, 0, 0, 1, 0, 1, 0},    // 3: ss, canonical
      {48, 32, 0, 1, 0, 1, 0, 1, 0},    // 4: ss, swapped
      {64, 112, 1, 0, 0, 1, 1, 4, 0},   // 5: ts, base_fc.0 shape, bias block, 4 groups
      {64, 112, 1, 0, 0, 3, 1, 2, 0},   // 6: ts, 3-pass, 2 groups
      {16, 48, 1, 0, 0, 3, 1, 2, 0},    // 7: ts, rgb_fc.0 shape
      {32, 64, 1, 0, 0, 1, 1, 4, 0},    // 8: ts, base_fc.2 shape
      {64, 112, 0, 0, 0, 3, 0, 1, 0},   // 9: ss 3-pass
      {48, 32, 1, 2, 0, 1, 0, 1, 0},    // 10: ts, descriptor fields exchanged
      {48, 32, 0, 2, 0, 1, 0, 1, 0},    // 11: ss, descriptor fields exchanged
      {48, 48, 1, 0, 0, 1, 1, 4, 0},    // 12: ts, vis_fc.2-like with 4 groups
      {48, 32, 1, 0, 0, 1, 0, 1, 1},    // 13: ts, B' = W^T MN-major (LBO=128, SBO=K*16)
      {48, 32, 1, 0, 0, 1, 0, 1, 2},    // 14: ts, B' MN-major, fields exchanged
      {64, 64, 1, 0, 0, 3, 0, 2, 1},    // 15: ts, 3-pass, MN-major, base_fc.0^T-like (first N half)
      {112, 64, 1, 0, 0, 3, 0, 2, 1},   // 16: ts, N = 112
  };
  const int nvar = sizeof(table) / sizeof(table[0]);
  if (variant < 0 || variant >= nvar) { printf("variant out of range\n"); return 2; }
  const Cfg c = table[variant];
  const int rows = 128 * c.groups;
  std::vector<float> A((size_t)rows * c.K), B((size_t)c.N * c.K), bias(c.N), D((size_t)rows * c.N, -777.f);
  srand(1234 + variant);
  auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (auto& x : A) x = rnd();
  for (auto& x : B) x = rnd() * 0.3f;
  for (auto& x : bias) x = rnd();
  float *dA, *dB, *dbias, *dD; int* dst;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dbias, bias.size() * 4);
  cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dst, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dbias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dst, 0, 4);
  const int KB = c.K + 16;
  size_t smem = (size_t)2 * c.N * KB * 2 + (c.a_in_tmem ? 0 : (size_t)c.groups * 2 * 128 * c.K * 2);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_probe<<<1, 128 * c.groups, smem>>>(c, dA, dB, dbias, dD, dst);
  cudaError_t e = cudaDeviceSynchronize();
  int st = 0;
  if (e == cudaSuccess) {
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
  }
  // references: exact fp32-input product (double) and bf16-rounded-input product
  double err_exact = 0, err_bf = 0, mag = 0;
  for (int r = 0; r < rows; ++r)
    for (int n = 0; n < c.N; ++n) {
      double se = 0, sb = 0;
      for (int k = 0; k < c.K; ++k) {
        se += (double)A[(size_t)r * c.K + k] * B[(size_t)n * c.K + k];
        sb += (double)bf16r(A[(size_t)r * c.K + k]) * bf16r(B[(size_t)n * c.K + k]);
      }
      if (c.ones_bias && c.a_in_tmem) { se += bias[n]; sb += bf16r(bias[n]); }
      const double d = D[(size_t)r * c.N + n];
      err_exact = fmax(err_exact, fabs(d - se));
      err_bf = fmax(err_bf, fabs(d - sb));
      mag = fmax(mag, fabs(se));
    }
  printf("variant %d: bmn=%d N=%d K=%d %s swap_lbo=%d swap_pack=%d passes=%d bias=%d groups=%d | cuda=%s timeout=%d | "
         "max|D-exact|=%.3e max|D-bf16ref|=%.3e (max|D|=%.3f) => %s\n",
         variant, c.b_mn, c.N, c.K, c.a_in_tmem ? "ts" : "ss", c.swap_lbo, c.swap_pack, c.passes, c.ones_bias, c.groups,
         cudaGetErrorString(e), st, err_exact, err_bf, mag,
         (e == cudaSuccess && !st && ((c.passes == 1 && err_bf < 1e-4) || (c.passes == 3 && err_exact < 1e-4))) ? "PASS" : "FAIL");
  return 0;
}
