// Shared helpers for the nerfool_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/nerfool_b200.h"

// ---------------------------------------------------------------------------------------------------
// error plumbing (nfb_api.cu owns the thread-local buffer)
// ---------------------------------------------------------------------------------------------------
int nfb_set_error(int code, const char* fmt, ...);

#define NFB_REQUIRE(cond, code, ...)                      \
  do {                                                    \
    if (!(cond)) return nfb_set_error((code), __VA_ARGS__); \
  } while (0)

#define NFB_CHECK_LAUNCH(name)                                                              \
  do {                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess)                                                                 \
      return nfb_set_error(NFB_ECUDA, "%s: launch failed: %s", (name), cudaGetErrorString(e__)); \
  } while (0)

int nfb_num_sms();

// First use of the library's device code in a process: resolve a kernel through cudaFuncGetAttributes before the first <<<>>>
// launch.  Launching first makes cudart probe cuKernelGetFunction on a module it has not loaded into the context yet; it
// recovers, but compute-sanitizer reports the probe as "CUDA_ERROR_INVALID_HANDLE ... cuKernelGetFunction" (the one API error of
// the round-1 memcheck log).  Used by the entry points that launch without a preceding cudaFuncSetAttribute.
#define NFB_RESOLVE_ONCE(kernel, who)                                                                      \
  do {                                                                                                     \
    static bool resolved__ = false;                                                                        \
    if (!resolved__) {                                                                                     \
      cudaFuncAttributes fa__;                                                                             \
      cudaError_t e__ = cudaFuncGetAttributes(&fa__, kernel);                                              \
      if (e__ != cudaSuccess) return nfb_set_error(NFB_ECUDA, "%s: cudaFuncGetAttributes: %s", (who), cudaGetErrorString(e__)); \
      resolved__ = true;                                                                                   \
    }                                                                                                      \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// IBRNet parameter blob layout (torch-native [out][in] tensors, mlp_network.py:153-208)
// ---------------------------------------------------------------------------------------------------
enum : int {
  P_S = 0,                                  // s (anti-alias temperature), 1
  P_DIR0_W = P_S + 1,                       // ray_dir_fc.0.weight [16][4]
  P_DIR0_B = P_DIR0_W + 16 * 4,             // [16]
  P_DIR2_W = P_DIR0_B + 16,                 // ray_dir_fc.2.weight [35][16]
  P_DIR2_B = P_DIR2_W + 35 * 16,            // [35]
  P_BASE0_W = P_DIR2_B + 35,                // base_fc.0.weight [64][105]
  P_BASE0_B = P_BASE0_W + 64 * 105,         // [64]
  P_BASE2_W = P_BASE0_B + 64,               // base_fc.2.weight [32][64]
  P_BASE2_B = P_BASE2_W + 32 * 64,          // [32]
  P_VIS0_W = P_BASE2_B + 32,                // vis_fc.0.weight [32][32]
  P_VIS0_B = P_VIS0_W + 32 * 32,            // [32]
  P_VIS2_W = P_VIS0_B + 32,                 // vis_fc.2.weight [33][32]
  P_VIS2_B = P_VIS2_W + 33 * 32,            // [33]
  P_VISB0_W = P_VIS2_B + 33,                // vis_fc2.0.weight [32][32]
  P_VISB0_B = P_VISB0_W + 32 * 32,          // [32]
  P_VISB2_W = P_VISB0_B + 32,               // vis_fc2.2.weight [1][32]
  P_VISB2_B = P_VISB2_W + 32,               // [1]
  P_GEO0_W = P_VISB2_B + 1,                 // geometry_fc.0.weight [64][65]
  P_GEO0_B = P_GEO0_W + 64 * 65,            // [64]
  P_GEO2_W = P_GEO0_B + 64,                 // geometry_fc.2.weight [16][64]
  P_GEO2_B = P_GEO2_W + 16 * 64,            // [16]
  P_ATT_Q = P_GEO2_B + 16,                  // ray_attention.w_qs.weight [16][16]
  P_ATT_K = P_ATT_Q + 256,
  P_ATT_V = P_ATT_K + 256,
  P_ATT_FC = P_ATT_V + 256,                 // ray_attention.fc.weight [16][16]
  P_LN_W = P_ATT_FC + 256,                  // ray_attention.layer_norm.weight [16]
  P_LN_B = P_LN_W + 16,
  P_OG0_W = P_LN_B + 16,                    // out_geometry_fc.0.weight [16][16]
  P_OG0_B = P_OG0_W + 256,
  P_OG2_W = P_OG0_B + 16,                   // out_geometry_fc.2.weight [1][16]
  P_OG2_B = P_OG2_W + 16,
  P_RGB0_W = P_OG2_B + 1,                   // rgb_fc.0.weight [16][37]
  P_RGB0_B = P_RGB0_W + 16 * 37,
  P_RGB2_W = P_RGB0_B + 16,                 // rgb_fc.2.weight [8][16]
  P_RGB2_B = P_RGB2_W + 8 * 16,
  P_RGB4_W = P_RGB2_B + 8,                  // rgb_fc.4.weight [1][8]
  P_RGB4_B = P_RGB4_W + 8,
  P_TOTAL = P_RGB4_B + 1
};
static_assert(P_TOTAL == NFB_IBRNET_PARAM_FLOATS, "IBRNet parameter count (mlp_network.py:153-208)");

// per-sample interface buffer slots
enum : int { PS_MEAN = 0, PS_VAR = 32, PS_WMEAN = 64, PS_RGB = 65, PS_NVALID = 68 };

// ---------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : (__expf(x) - 1.f); }
// derivative of ELU expressed through its OUTPUT y (torch's in-place ELU backward uses the result)
__device__ __forceinline__ float elu_grad_from_out(float y) { return y > 0.f ? 1.f : (y + 1.f); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
