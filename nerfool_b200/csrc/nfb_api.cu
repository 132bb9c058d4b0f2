// Library-level entry points: version, thread-local error text, parameter-blob layout lookup.
#include <stdarg.h>
#include <string.h>
#include "nfb_common.cuh"

static thread_local char g_err[512] = "";

int nfb_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int nfb_num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

extern "C" int nfb_version(void) { return 100; }
extern "C" const char* nfb_last_error_string(void) { return g_err; }

extern "C" int nfb_ibrnet_param_offset(const char* name) {
  static const struct { const char* n; int off; } tab[] = {
      {"s", P_S},
      {"ray_dir_fc.0.weight", P_DIR0_W}, {"ray_dir_fc.0.bias", P_DIR0_B},
      {"ray_dir_fc.2.weight", P_DIR2_W}, {"ray_dir_fc.2.bias", P_DIR2_B},
      {"base_fc.0.weight", P_BASE0_W}, {"base_fc.0.bias", P_BASE0_B},
      {"base_fc.2.weight", P_BASE2_W}, {"base_fc.2.bias", P_BASE2_B},
      {"vis_fc.0.weight", P_VIS0_W}, {"vis_fc.0.bias", P_VIS0_B},
      {"vis_fc.2.weight", P_VIS2_W}, {"vis_fc.2.bias", P_VIS2_B},
      {"vis_fc2.0.weight", P_VISB0_W}, {"vis_fc2.0.bias", P_VISB0_B},
      {"vis_fc2.2.weight", P_VISB2_W}, {"vis_fc2.2.bias", P_VISB2_B},
      {"geometry_fc.0.weight", P_GEO0_W}, {"geometry_fc.0.bias", P_GEO0_B},
      {"geometry_fc.2.weight", P_GEO2_W}, {"geometry_fc.2.bias", P_GEO2_B},
      {"ray_attention.w_qs.weight", P_ATT_Q}, {"ray_attention.w_ks.weight", P_ATT_K},
      {"ray_attention.w_vs.weight", P_ATT_V}, {"ray_attention.fc.weight", P_ATT_FC},
      {"ray_attention.layer_norm.weight", P_LN_W}, {"ray_attention.layer_norm.bias", P_LN_B},
      {"out_geometry_fc.0.weight", P_OG0_W}, {"out_geometry_fc.0.bias", P_OG0_B},
      {"out_geometry_fc.2.weight", P_OG2_W}, {"out_geometry_fc.2.bias", P_OG2_B},
      {"rgb_fc.0.weight", P_RGB0_W}, {"rgb_fc.0.bias", P_RGB0_B},
      {"rgb_fc.2.weight", P_RGB2_W}, {"rgb_fc.2.bias", P_RGB2_B},
      {"rgb_fc.4.weight", P_RGB4_W}, {"rgb_fc.4.bias", P_RGB4_B},
  };
  if (!name) return -1;
  for (size_t i = 0; i < sizeof(tab) / sizeof(tab[0]); ++i)
    if (strcmp(tab[i].n, name) == 0) return tab[i].off;
  return -1;
}
