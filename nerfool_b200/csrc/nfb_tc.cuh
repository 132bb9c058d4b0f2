// tcgen05 / TMEM / mbarrier primitives (sm_100a inline PTX) used by the tensor-core IBRNet kernels.
//
// Operand conventions used throughout this repo
//   * D (fp32 accumulator) lives in TMEM: row i of a 128-row tile = TMEM lane i, output n = column d_col + n.
//   * A (bf16) lives in TMEM as well (the ".ts" form of tcgen05.mma): row i = lane i, elements (2j, 2j+1) of the
//     row packed in the (low, high) halves of column a_col + j.  The thread that owns row i writes it with
//     tcgen05.st.32x32b and reads D back with tcgen05.ld.32x32b -- no shared-memory staging, no swizzle.
//   * B (bf16 layer weights, torch layout [N][K] = "K-major") lives in shared memory in the canonical
//     no-swizzle K-major layout of the UMMA shared-memory descriptor: 8x8 "core matrices" of 8 rows x 16 bytes
//     (128 contiguous bytes); element (n, k) sits at
//         (k / 8) * LBO + (n / 8) * SBO + (n % 8) * 16 + (k % 8) * 2        bytes,
//     LBO = "leading byte offset" = distance between core matrices adjacent in K,
//     SBO = "stride byte offset"  = distance between core matrices adjacent in N.
//     We store tiles K-chunk-major: SBO = 128, LBO = N * 16, so that a K=16 slice (one MMA) is 2*LBO bytes.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nfbtc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------------------
// TMEM allocation (one warp, .sync.aligned) -- ncols power of two in [32, 512]
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// potentially-blocking test: the hardware suspends the thread until the phase completes or the time hint (ns)
// expires, so the retry loop around it costs almost no issue slots
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// completion of all previously issued tcgen05.mma of this thread -> one arrival on the mbarrier
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// descriptors
// ---------------------------------------------------------------------------------------------------
// instruction descriptor for kind::f16 with bf16 A/B, fp32 D, both operands K-major, M x N tile
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4)                      // D format: F32
         | (1u << 7)                    // A format: BF16
         | (1u << 10)                   // B format: BF16
         | ((uint32_t)(N >> 3) << 17)   // N / 8
         | ((uint32_t)(M >> 4) << 24);  // M / 16
}
// shared-memory matrix descriptor: no swizzle, K-major, version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// D[tmem] (+)= A[tmem] * B[smem]^T ; one K=16 step ; issued by ONE thread
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// ---------------------------------------------------------------------------------------------------
// TMEM <-> registers, 32x32b shape: lane i of the warp <-> TMEM lane (taddr.lane + i), N consecutive columns
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
// raw 32-bit words
__device__ __forceinline__ void tmem_ld16u(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8u(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld4u(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}

// ELU-derivative stash.  For y = ELU(x): ELU'(x) = 1 (y > 0) or y + 1 (y <= 0), i.e. ELU' = 2 - t with
// t = 1 - min(y, 0) in [1, 2].  t is clamped just below 2 and bits [23:8] of its fp32 pattern (15 mantissa bits,
// truncation error < 2^-15) are kept, two values per 32-bit word, so that encode and decode are one PRMT each.
__device__ __forceinline__ uint32_t elu_stash_pack(float y0, float y1) {
  const float t0 = fminf(1.f - fminf(y0, 0.f), 1.9999999f);
  const float t1 = fminf(1.f - fminf(y1, 0.f), 1.9999999f);
  return __byte_perm(__float_as_uint(t0), __float_as_uint(t1), 0x6521);   // {t0.b1, t0.b2, t1.b1, t1.b2}
}
__device__ __forceinline__ float elu_stash_lo(uint32_t w) {   // ELU' of the first value of the pair
  return 2.f - __uint_as_float(__byte_perm(w, 0x3F000000u, 0x7104));      // {0, w.b0, w.b1, 0x3F}
}
__device__ __forceinline__ float elu_stash_hi(uint32_t w) {
  return 2.f - __uint_as_float(__byte_perm(w, 0x3F000000u, 0x7324));      // {0, w.b2, w.b3, 0x3F}
}

// ELU with one MUFU and no branch: elu(x) = max(x, min(exp(x) - 1, 0))
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float elu_fast(float x) {
  return fmaxf(x, fminf(ex2_approx(x * 1.4426950408889634f) - 1.f, 0.f));
}
__device__ __forceinline__ float sigmoid_fast(float x) {
  return __fdividef(1.f, 1.f + ex2_approx(-1.4426950408889634f * x));
}

// two floats -> packed bf16x2 (round to nearest even): low half = a, high half = b
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// split (a, b) into hi = bf16(x) and lo = bf16(x - hi): x = hi + lo to ~2^-17 relative
// (the residual pair is one packed FFMA2: on sm_100 scalar fp32 instructions issue at half the packed rate)
__device__ __forceinline__ void split_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(a, b);
  const float2 h = make_float2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xFFFF0000u));
  const float2 l = __ffma2_rn(h, make_float2(-1.f, -1.f), make_float2(a, b));
  lo = pack_bf16(l.x, l.y);
}

// ---- packed (two-lane) forms of the epilogue arithmetic ----
__device__ __forceinline__ float2 elu_fast2(float2 x) {
  const float2 t = __fmul2_rn(x, make_float2(1.4426950408889634f, 1.4426950408889634f));
  const float2 em = __fadd2_rn(make_float2(ex2_approx(t.x), ex2_approx(t.y)), make_float2(-1.f, -1.f));
  return make_float2(fmaxf(x.x, fminf(em.x, 0.f)), fmaxf(x.y, fminf(em.y, 0.f)));
}
// ELU of a pair and the 16-bit derivative codes of both.  The exponent argument is clamped to <= 0, so e = exp(min(x, 0)) lies in
// (0, 1]: y = max(x, e - 1) is ELU for either sign, ELU' = e needs no clamp, and the code word t = 1.9999999 - e (1 - 2^-22)
// stays inside [1, 2) by construction (4 FMNMX per pair instead of 6).
__device__ __forceinline__ float2 elu_code2(float2 x, uint32_t& code) {
  float2 t = __fmul2_rn(x, make_float2(1.4426950408889634f, 1.4426950408889634f));
  t.x = fminf(t.x, 0.f); t.y = fminf(t.y, 0.f);
  const float2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));
  const float2 em = __fadd2_rn(e, make_float2(-1.f, -1.f));
  const float2 c = __ffma2_rn(e, make_float2(-0.99999976f, -0.99999976f), make_float2(1.9999999f, 1.9999999f));
  code = __byte_perm(__float_as_uint(c.x), __float_as_uint(c.y), 0x6521);
  return make_float2(fmaxf(x.x, em.x), fmaxf(x.y, em.y));
}
// derivative pair of a code word: (ELU'_lo, ELU'_hi) = 2 - (t_lo, t_hi)
__device__ __forceinline__ float2 elu_stash2(uint32_t w) {
  const float2 t = make_float2(__uint_as_float(__byte_perm(w, 0x3F000000u, 0x7104)), __uint_as_float(__byte_perm(w, 0x3F000000u, 0x7324)));
  return __ffma2_rn(t, make_float2(-1.f, -1.f), make_float2(2.f, 2.f));
}
// (a, b) *= ELU' pair
__device__ __forceinline__ void mul_stash2(float& a, float& b, uint32_t w) {
  const float2 r = __fmul2_rn(make_float2(a, b), elu_stash2(w));
  a = r.x; b = r.y;
}

}  // namespace nfbtc
