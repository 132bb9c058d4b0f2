// forward_warp (eval/ibrnet/eval_adv.py:97-197): z-buffer splat of one view's pixels into another view, the step the
// depth-consistency / camera-consistency attack losses run once per PGD iteration.  The reference does it with a Python loop
// over all H*W pixels on CPU tensors; here it is two streaming passes over the pixels (HBM-bound integer work).
//
// Reference semantics (sequential over the source pixels i in list order):
//     if new_depth[dst_i] == 0 or new_depth[dst_i] > depth_i:  new_depth[dst_i] = depth_i ; new_rgb[dst_i] = rgb_i
// i.e. for strictly positive depths the destination keeps the MINIMUM depth and, among equal minima, the FIRST source (strict
// '>').  That is an order-independent reduction: pass 1 takes atomicMin over 64-bit keys (depth bits << 32 | list position),
// pass 2 resolves the winners.  A depth <= 0 (or NaN) breaks the equivalence (0 doubles as the "empty" marker): pass 1 raises
// a flag and the host wrapper re-runs the exact sequential loop in one thread (k_warp_sequential) -- a correctness fallback.
#include "nfb_common.cuh"

namespace {
constexpr unsigned long long WARP_EMPTY = ~0ull;

__global__ void k_warp_init(int n, unsigned long long* keys, float* new_rgb, float* new_depth, int* flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *flag = 0;
  if (i < n) {
    keys[i] = WARP_EMPTY;
    new_depth[i] = 0.f;
    new_rgb[3 * i] = 0.f; new_rgb[3 * i + 1] = 0.f; new_rgb[3 * i + 2] = 0.f;
  }
}

__global__ void k_warp_min(int W, int n_src, const int* __restrict__ sources, const int* __restrict__ x_res,
                           const int* __restrict__ y_res, const float* __restrict__ depth_src,
                           const uint8_t* __restrict__ allowed, unsigned long long* keys, int* flag) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n_src) return;
  const int i = sources ? __ldg(sources + pos) : pos;
  const int dst = __ldg(y_res + i) * W + __ldg(x_res + i);
  if (allowed && !__ldg(allowed + dst)) return;
  const float d = __ldg(depth_src + i);
  if (!(d > 0.f)) {                       // <= 0 or NaN: the order-independent form does not apply
    atomicOr(flag, 1);
    return;
  }
  const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)pos;
  atomicMin(keys + dst, key);
}

__global__ void k_warp_resolve(int n, const int* __restrict__ sources, const unsigned long long* __restrict__ keys,
                               const float* __restrict__ depth_src, const float* __restrict__ rgb_ref, float* new_rgb,
                               float* new_depth) {
  const int dst = blockIdx.x * blockDim.x + threadIdx.x;
  if (dst >= n) return;
  const unsigned long long key = keys[dst];
  if (key == WARP_EMPTY) return;
  const int pos = (int)(key & 0xffffffffu);
  const int i = sources ? __ldg(sources + pos) : pos;
  new_depth[dst] = __ldg(depth_src + i);
  new_rgb[3 * dst] = __ldg(rgb_ref + 3 * i);
  new_rgb[3 * dst + 1] = __ldg(rgb_ref + 3 * i + 1);
  new_rgb[3 * dst + 2] = __ldg(rgb_ref + 3 * i + 2);
}

// the reference loop, verbatim, in one thread (only when a depth <= 0 was seen)
__global__ void k_warp_sequential(int n, int W, int n_src, const int* __restrict__ sources, const int* __restrict__ x_res,
                                  const int* __restrict__ y_res, const float* __restrict__ depth_src,
                                  const float* __restrict__ rgb_ref, const uint8_t* __restrict__ allowed, float* new_rgb,
                                  float* new_depth) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int k = 0; k < n; ++k) { new_depth[k] = 0.f; new_rgb[3 * k] = new_rgb[3 * k + 1] = new_rgb[3 * k + 2] = 0.f; }
  for (int pos = 0; pos < n_src; ++pos) {
    const int i = sources ? sources[pos] : pos;
    const int dst = y_res[i] * W + x_res[i];
    if (allowed && !allowed[dst]) continue;
    const float d = depth_src[i];
    if (new_depth[dst] == 0.f || new_depth[dst] > d) {
      new_depth[dst] = d;
      new_rgb[3 * dst] = rgb_ref[3 * i]; new_rgb[3 * dst + 1] = rgb_ref[3 * i + 1]; new_rgb[3 * dst + 2] = rgb_ref[3 * i + 2];
    }
  }
}
}  // namespace

extern "C" int nfb_forward_warp(int H, int W, const int* x_res, const int* y_res, const float* depth_src, const float* rgb_ref,
                                const uint8_t* allowed, const int* sources, int n_src, float* new_rgb, float* new_depth,
                                unsigned long long* keys, int* flag, int sequential, void* stream) {
  NFB_REQUIRE(H >= 1 && W >= 1 && n_src >= 0, NFB_EINVAL, "nfb_forward_warp: bad arguments (H=%d W=%d n_src=%d)", H, W, n_src);
  NFB_REQUIRE(x_res && y_res && depth_src && rgb_ref && new_rgb && new_depth && keys && flag, NFB_EINVAL, "nfb_forward_warp: NULL buffer");
  NFB_REQUIRE((long long)H * W < (1ll << 31), NFB_EUNSUPPORTED, "nfb_forward_warp: image too large");
  const int n = H * W;
  cudaStream_t st = (cudaStream_t)stream;
  const int ns = sources ? n_src : n;
  NFB_RESOLVE_ONCE(k_warp_init, "nfb_forward_warp");
  if (sequential) {
    k_warp_sequential<<<1, 32, 0, st>>>(n, W, ns, sources, x_res, y_res, depth_src, rgb_ref, allowed, new_rgb, new_depth);
    NFB_CHECK_LAUNCH("k_warp_sequential");
    return NFB_OK;
  }
  const int T = 256;
  k_warp_init<<<(n + T - 1) / T, T, 0, st>>>(n, keys, new_rgb, new_depth, flag);
  NFB_CHECK_LAUNCH("k_warp_init");
  if (ns > 0) {
    k_warp_min<<<(ns + T - 1) / T, T, 0, st>>>(W, ns, sources, x_res, y_res, depth_src, allowed, keys, flag);
    NFB_CHECK_LAUNCH("k_warp_min");
  }
  k_warp_resolve<<<(n + T - 1) / T, T, 0, st>>>(n, sources, keys, depth_src, rgb_ref, new_rgb, new_depth);
  NFB_CHECK_LAUNCH("k_warp_resolve");
  return NFB_OK;
}
