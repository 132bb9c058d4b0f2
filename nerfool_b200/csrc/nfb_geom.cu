// Bandwidth-side kernels of the hot path: depth sampling, projection + bilinear gather (and its scatter
// backward), alpha compositing (and backward), inverse-CDF importance sampling + merge.
#include "nfb_geom.cuh"

// =====================================================================================================
// sample_along_camera_ray (render_ray.py:73-116)
// =====================================================================================================
// Every operation is a separately rounded IEEE op in the reference's order so z is bit-identical:
//   inv_uniform: start = 1/near; step = (1/far - start)/(S-1); z_i = 1/(start + i*step)
//   else       : step = (far-near)/(S-1);                      z_i = near + i*step
__device__ __forceinline__ float coarse_z_at(int i, float near_d, float far_d, int S, int inv_uniform) {
  if (inv_uniform) {
    const float start = __fdiv_rn(1.f, near_d);
    const float step = __fdiv_rn(__fsub_rn(__fdiv_rn(1.f, far_d), start), (float)(S - 1));
    return __fdiv_rn(1.f, __fadd_rn(start, __fmul_rn((float)i, step)));
  }
  const float step = __fdiv_rn(__fsub_rn(far_d, near_d), (float)(S - 1));
  return __fadd_rn(near_d, __fmul_rn((float)i, step));
}

__global__ void k_coarse_depths(int R, int S, float near_d, float far_d, int inv_uniform,
                                const float* __restrict__ t_rand, float* __restrict__ z_out) {
  const size_t n = (size_t)R * S;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % S);
    float z = coarse_z_at(i, near_d, far_d, S, inv_uniform);
    if (t_rand) {
      // stratified jitter (render_ray.py:104-111): lower/upper are the mid-points with the end samples kept
      const float zl = (i > 0) ? coarse_z_at(i - 1, near_d, far_d, S, inv_uniform) : z;
      const float zu = (i < S - 1) ? coarse_z_at(i + 1, near_d, far_d, S, inv_uniform) : z;
      const float lower = (i > 0) ? __fmul_rn(0.5f, __fadd_rn(z, zl)) : z;
      const float upper = (i < S - 1) ? __fmul_rn(0.5f, __fadd_rn(zu, z)) : z;
      z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand[idx]));
    }
    z_out[idx] = z;
  }
}

extern "C" int nfb_coarse_depths(int R, int S, float near_depth, float far_depth, int inv_uniform,
                                 const float* t_rand, float* z_out, void* stream) {
  NFB_REQUIRE(R >= 0 && S >= 2, NFB_EINVAL, "nfb_coarse_depths: bad arguments (R=%d S=%d)", R, S);
  NFB_REQUIRE(near_depth > 0 && far_depth > near_depth, NFB_EINVAL,
              "nfb_coarse_depths: need 0 < near < far (render_ray.py:87)");
  if (R == 0) return NFB_OK;   // empty ray batch: nothing to do, pointers may be NULL
  NFB_REQUIRE(z_out, NFB_EINVAL, "nfb_coarse_depths: bad arguments (z_out is NULL)");
  NFB_RESOLVE_ONCE(k_coarse_depths, "nfb_coarse_depths");
  const size_t n = (size_t)R * S;
  const int block = 256;
  const int grid = (int)((n + block - 1) / block < (size_t)nfb_num_sms() * 8 ? (n + block - 1) / block
                                                                            : (size_t)nfb_num_sms() * 8);
  k_coarse_depths<<<grid, block, 0, (cudaStream_t)stream>>>(R, S, near_depth, far_depth, inv_uniform, t_rand, z_out);
  NFB_CHECK_LAUNCH("k_coarse_depths");
  return NFB_OK;
}

// =====================================================================================================
// Projector.compute (projection.py:89-132): one thread per (point, view) row, 128 rows per tile.
// The 35-float rows of a tile are contiguous in rgb_feat, so they are staged in shared memory
// ([row][35], odd stride = conflict free) and written back as one coalesced stream.
// =====================================================================================================
constexpr int PG_ROWS = 128;

__global__ void __launch_bounds__(PG_ROWS)
k_project_gather_fwd(int N, int V, int H, int W, int fh, int fw, PointSrc psrc, const float* __restrict__ cam,
                     const float* __restrict__ imgs, const float* __restrict__ feat,
                     float* __restrict__ rgb_feat, float* __restrict__ ray_diff, float* __restrict__ mask) {
  __shared__ float stage[PG_ROWS * NFB_ROW_CH];
  __shared__ __align__(16) float coop[PG_ROWS * 36];       // transposition rows of the cooperative gather
  __shared__ float s_cam[16 * NFB_MAX_VIEWS + 4];
  for (int i = threadIdx.x; i < 16 * V + 3; i += blockDim.x) s_cam[i] = cam[i];
  __syncthreads();
  const float Wm1 = (float)W - 1.f, Hm1 = (float)H - 1.f;
  const size_t total = (size_t)N * V;
  const size_t ntiles = (total + PG_ROWS - 1) / PG_ROWS;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t row0 = tile * PG_ROWS;
    const size_t row = row0 + threadIdx.x;
    {
      // quarter-warp cooperative gather (gather_row_coop): every lane of the warp takes part, rows past the end are inactive
      const bool act = row < total;
      const int p = act ? (int)(row / V) : 0, v = act ? (int)(row % V) : 0;
      ViewGeom g;
      g.gx = g.gy = g.mask = 0.f; g.rd[0] = g.rd[1] = g.rd[2] = g.rd[3] = 0.f;
      if (act) {
        float x, y, z;
        load_point(psrc, p, x, y, z);
        g = view_geometry(x, y, z, s_cam + 16 * v, s_cam + 16 * V, Wm1, Hm1);
      }
      float r[NFB_ROW_CH];
      gather_row_coop(act, g, v, H, W, fh, fw, imgs, feat, coop, 36, (int)threadIdx.x, r);
      if (act) {
#pragma unroll
        for (int c = 0; c < NFB_ROW_CH; ++c) stage[threadIdx.x * NFB_ROW_CH + c] = r[c];
        reinterpret_cast<float4*>(ray_diff)[row] = make_float4(g.rd[0], g.rd[1], g.rd[2], g.rd[3]);
        mask[row] = g.mask;
      }
    }
    __syncthreads();
    const size_t rows_here = (total - row0 < (size_t)PG_ROWS) ? (total - row0) : (size_t)PG_ROWS;
    const int nfl = (int)rows_here * NFB_ROW_CH;
    float* dst = rgb_feat + row0 * NFB_ROW_CH;
    for (int i = threadIdx.x; i < nfl; i += blockDim.x) dst[i] = stage[i];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(PG_ROWS)
k_project_gather_bwd(int N, int V, int H, int W, int fh, int fw, PointSrc psrc, const float* __restrict__ cam,
                     const float* __restrict__ d_rgb_feat, float* __restrict__ d_feat,
                     float* __restrict__ d_imgs) {
  __shared__ float stage[PG_ROWS * NFB_ROW_CH];
  __shared__ float s_cam[16 * NFB_MAX_VIEWS + 4];
  for (int i = threadIdx.x; i < 16 * V + 3; i += blockDim.x) s_cam[i] = cam[i];
  __syncthreads();
  const float Wm1 = (float)W - 1.f, Hm1 = (float)H - 1.f;
  const size_t total = (size_t)N * V;
  const size_t ntiles = (total + PG_ROWS - 1) / PG_ROWS;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t row0 = tile * PG_ROWS;
    const size_t rows_here = (total - row0 < (size_t)PG_ROWS) ? (total - row0) : (size_t)PG_ROWS;
    const int nfl = (int)rows_here * NFB_ROW_CH;
    const float* src = d_rgb_feat + row0 * NFB_ROW_CH;
    for (int i = threadIdx.x; i < nfl; i += blockDim.x) stage[i] = __ldg(src + i);
    __syncthreads();
    const size_t row = row0 + threadIdx.x;
    if (row < total) {
      const int p = (int)(row / V), v = (int)(row % V);
      float x, y, z;
      load_point(psrc, p, x, y, z);
      const ViewGeom g = view_geometry(x, y, z, s_cam + 16 * v, s_cam + 16 * V, Wm1, Hm1);
      float r[NFB_ROW_CH];
#pragma unroll
      for (int c = 0; c < NFB_ROW_CH; ++c) r[c] = stage[threadIdx.x * NFB_ROW_CH + c];
      scatter_row(g, v, H, W, fh, fw, r, d_feat, d_imgs);
    }
    __syncthreads();
  }
}

static int check_geometry_args(const char* who, int N, int S, int V, int H, int W, int fh, int fw,
                               const float* xyz, const float* ray_o, const float* ray_d, const float* z,
                               const float* cam) {
  NFB_REQUIRE(N >= 0 && V >= 1 && H >= 2 && W >= 2 && fh >= 1 && fw >= 1, NFB_EINVAL,
              "%s: bad sizes N=%d V=%d H=%d W=%d fh=%d fw=%d", who, N, V, H, W, fh, fw);
  NFB_REQUIRE(V <= NFB_MAX_VIEWS, NFB_EUNSUPPORTED, "%s: V=%d > %d views", who, V, NFB_MAX_VIEWS);
  if (N == 0) return NFB_OK;
  NFB_REQUIRE(cam != nullptr, NFB_EINVAL, "%s: cam is NULL", who);
  if (!xyz) {
    NFB_REQUIRE(ray_o && ray_d && z && S >= 1 && (N % S) == 0, NFB_EINVAL,
                "%s: implicit points need ray_o, ray_d, z and S | N (N=%d S=%d)", who, N, S);
  }
  return NFB_OK;
}

static inline int tiles_grid(size_t rows, int rows_per_tile, int ctas_per_sm) {
  const size_t tiles = (rows + rows_per_tile - 1) / rows_per_tile;
  const size_t cap = (size_t)nfb_num_sms() * ctas_per_sm;
  return (int)(tiles < cap ? (tiles ? tiles : 1) : cap);
}

extern "C" int nfb_project_gather_fwd(int N, int S, int V, int H, int W, int fh, int fw,
                                      const float* xyz, const float* ray_o, const float* ray_d, const float* z,
                                      const float* cam, const float* imgs, const float* feat,
                                      float* rgb_feat, float* ray_diff, float* mask, void* stream) {
  int rc = check_geometry_args("nfb_project_gather_fwd", N, S, V, H, W, fh, fw, xyz, ray_o, ray_d, z, cam);
  if (rc) return rc;
  if (N == 0) return NFB_OK;
  NFB_REQUIRE(imgs && feat && rgb_feat && ray_diff && mask, NFB_EINVAL, "nfb_project_gather_fwd: NULL buffer");
  NFB_REQUIRE(((uintptr_t)feat % 16) == 0 && ((uintptr_t)ray_diff % 16) == 0, NFB_EINVAL,
              "nfb_project_gather_fwd: feat / ray_diff must be 16-byte aligned");
  if (N == 0) return NFB_OK;
  NFB_RESOLVE_ONCE(k_project_gather_fwd, "nfb_project_gather_fwd");
  PointSrc ps{xyz, ray_o, ray_d, z, S};
  k_project_gather_fwd<<<tiles_grid((size_t)N * V, PG_ROWS, 8), PG_ROWS, 0, (cudaStream_t)stream>>>(
      N, V, H, W, fh, fw, ps, cam, imgs, feat, rgb_feat, ray_diff, mask);
  NFB_CHECK_LAUNCH("k_project_gather_fwd");
  return NFB_OK;
}

extern "C" int nfb_project_gather_bwd(int N, int S, int V, int H, int W, int fh, int fw,
                                      const float* xyz, const float* ray_o, const float* ray_d, const float* z,
                                      const float* cam, const float* d_rgb_feat, float* d_feat, float* d_imgs,
                                      void* stream) {
  int rc = check_geometry_args("nfb_project_gather_bwd", N, S, V, H, W, fh, fw, xyz, ray_o, ray_d, z, cam);
  if (rc) return rc;
  if (N == 0) return NFB_OK;
  NFB_REQUIRE(d_rgb_feat, NFB_EINVAL, "nfb_project_gather_bwd: d_rgb_feat is NULL");
  NFB_REQUIRE(((uintptr_t)d_feat % 16) == 0, NFB_EINVAL, "nfb_project_gather_bwd: d_feat must be 16-byte aligned");
  if (N == 0 || (!d_feat && !d_imgs)) return NFB_OK;
  PointSrc ps{xyz, ray_o, ray_d, z, S};
  k_project_gather_bwd<<<tiles_grid((size_t)N * V, PG_ROWS, 8), PG_ROWS, 0, (cudaStream_t)stream>>>(
      N, V, H, W, fh, fw, ps, cam, d_rgb_feat, d_feat, d_imgs);
  NFB_CHECK_LAUNCH("k_project_gather_bwd");
  return NFB_OK;
}

// ---------------------------------------------------------------------------------------------------
// grid_sampler_2d backward w.r.t. the GRID, per (point, view) row: d loss / d (gx, gy) of both bilinear gathers.
// gnt/projection.py:84-132 does not detach the source cameras, so eval/gnt/eval_adv.py --perturb_camera differentiates the
// sampling positions; the chain from (gx, gy) to the 34-float camera vectors is a handful of torch ops on the host side
// (ops.ProjectGatherCam).  Same tap set / weights as the forward (ATen: t = x - floor(x), zero padding per tap).
// ---------------------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ void grid_grad_accum(float gx, float gy, int w, int h, const float* __restrict__ base /*[h][w][C]*/,
                                                const float* __restrict__ g /*[C]*/, float& dgx, float& dgy) {
  const float ix = __fmul_rn(__fadd_rn(gx, 1.f), 0.5f * (float)(w - 1));
  const float iy = __fmul_rn(__fadd_rn(gy, 1.f), 0.5f * (float)(h - 1));
  const float fx = floorf(ix), fy = floorf(iy);
  const float tw = ix - fx, te = 1.f - tw, tn = iy - fy, ts = 1.f - tn;
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const bool xw = (x0 >= 0) && (x0 < w), xe = (x1 >= 0) && (x1 < w);
  const bool yn = (y0 >= 0) && (y0 < h), ys = (y1 >= 0) && (y1 < h);
  const float* pnw = (xw && yn) ? base + ((size_t)y0 * w + x0) * C : nullptr;
  const float* pne = (xe && yn) ? base + ((size_t)y0 * w + x1) * C : nullptr;
  const float* psw = (xw && ys) ? base + ((size_t)y1 * w + x0) * C : nullptr;
  const float* pse = (xe && ys) ? base + ((size_t)y1 * w + x1) * C : nullptr;
  float ax = 0.f, ay = 0.f;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float nw = pnw ? __ldg(pnw + c) : 0.f, ne = pne ? __ldg(pne + c) : 0.f;
    const float sw = psw ? __ldg(psw + c) : 0.f, se = pse ? __ldg(pse + c) : 0.f;
    ax = fmaf(g[c], ts * (ne - nw) + tn * (se - sw), ax);
    ay = fmaf(g[c], te * (sw - nw) + tw * (se - ne), ay);
  }
  dgx = fmaf(ax, 0.5f * (float)(w - 1), dgx);
  dgy = fmaf(ay, 0.5f * (float)(h - 1), dgy);
}

__global__ void __launch_bounds__(128)
k_project_grid_bwd(int N, int V, int H, int W, int fh, int fw, PointSrc psrc, const float* __restrict__ cam,
                   const float* __restrict__ imgs, const float* __restrict__ feat, const float* __restrict__ d_rgb_feat,
                   float* __restrict__ d_grid) {
  __shared__ float s_cam[16 * NFB_MAX_VIEWS + 4];
  for (int i = threadIdx.x; i < 16 * V + 3; i += blockDim.x) s_cam[i] = cam[i];
  __syncthreads();
  const float Wm1 = (float)W - 1.f, Hm1 = (float)H - 1.f;
  const size_t total = (size_t)N * V;
  for (size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x; row < total; row += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(row / V), v = (int)(row % V);
    float x, y, z;
    load_point(psrc, p, x, y, z);
    const ViewGeom g = view_geometry(x, y, z, s_cam + 16 * v, s_cam + 16 * V, Wm1, Hm1);
    const float* gr = d_rgb_feat + row * NFB_ROW_CH;
    float dgx = 0.f, dgy = 0.f;
    grid_grad_accum<3>(g.gx, g.gy, W, H, imgs + (size_t)v * H * W * 3, gr, dgx, dgy);
    grid_grad_accum<NFB_FEAT_CH>(g.gx, g.gy, fw, fh, feat + (size_t)v * fh * fw * NFB_FEAT_CH, gr + 3, dgx, dgy);
    d_grid[row * 2] = dgx;
    d_grid[row * 2 + 1] = dgy;
  }
}

extern "C" int nfb_project_grid_bwd(int N, int S, int V, int H, int W, int fh, int fw,
                                    const float* xyz, const float* ray_o, const float* ray_d, const float* z,
                                    const float* cam, const float* imgs, const float* feat, const float* d_rgb_feat,
                                    float* d_grid, void* stream) {
  int rc = check_geometry_args("nfb_project_grid_bwd", N, S, V, H, W, fh, fw, xyz, ray_o, ray_d, z, cam);
  if (rc) return rc;
  if (N == 0) return NFB_OK;
  NFB_REQUIRE(imgs && feat && d_rgb_feat && d_grid, NFB_EINVAL, "nfb_project_grid_bwd: NULL buffer");
  PointSrc ps{xyz, ray_o, ray_d, z, S};
  k_project_grid_bwd<<<tiles_grid((size_t)N * V, 128, 16), 128, 0, (cudaStream_t)stream>>>(N, V, H, W, fh, fw, ps, cam, imgs, feat,
                                                                                        d_rgb_feat, d_grid);
  NFB_CHECK_LAUNCH("k_project_grid_bwd");
  return NFB_OK;
}

// =====================================================================================================
// raw2outputs (render_ray.py:123-170): one warp per ray, lanes stride the samples (coalesced), the
// transmittance is a warp scan.  The running product is kept in fp64 like torch's CPU cumprod.
// =====================================================================================================
__device__ __forceinline__ double warp_incl_prod(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= n;
  }
  return v;
}
__device__ __forceinline__ float warp_suffix_excl_sum(float v, int lane, float& total) {
  // returns sum over lanes > lane; total = sum over all lanes
  float s = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_down_sync(0xffffffffu, s, o);
    if (lane + o < 32) s += n;
  }
  total = __shfl_sync(0xffffffffu, s, 0);
  return s - v;
}

constexpr int CMP_WARPS = 4;

__global__ void __launch_bounds__(CMP_WARPS * 32)
k_composite_fwd(int R, int S, int white_bkgd, const float* __restrict__ raw, const float* __restrict__ z,
                const uint8_t* __restrict__ pixel_mask, const float* __restrict__ n_valid, int nv_stride,
                float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ weights,
                float* __restrict__ alpha, uint8_t* __restrict__ ray_mask) {
  const int lane = threadIdx.x & 31;
  const int wglobal = blockIdx.x * CMP_WARPS + (threadIdx.x >> 5);
  const int wstride = gridDim.x * CMP_WARPS;
  for (int r = wglobal; r < R; r += wstride) {
    double carry = 1.0;          // product of (1 - alpha + 1e-10) over all previous chunks
    float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_d = 0.f, acc_w = 0.f;
    int cnt = 0;
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      const bool in = s < S;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      float zz = 0.f;
      if (in) {
        q = __ldg(reinterpret_cast<const float4*>(raw) + (size_t)r * S + s);
        zz = __ldg(z + (size_t)r * S + s);
        const bool pm = pixel_mask ? (pixel_mask[(size_t)r * S + s] != 0)
                                   : (__ldg(n_valid + ((size_t)r * S + s) * nv_stride) > 1.f);
        cnt += pm ? 1 : 0;
      }
      const float a = in ? (1.f - expf(-q.w)) : 0.f;
      const float f = in ? __fadd_rn(__fsub_rn(1.f, a), 1e-10f) : 1.f;
      const double incl = warp_incl_prod((double)f, lane) * carry;
      double excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = carry;
      carry = __shfl_sync(0xffffffffu, incl, 31);
      const float T = (float)excl;
      const float w = a * T;
      if (in) {
        weights[(size_t)r * S + s] = w;
        alpha[(size_t)r * S + s] = a;
        acc_r += w * q.x; acc_g += w * q.y; acc_b += w * q.z; acc_d += w * zz; acc_w += w;
      }
    }
    acc_r = warp_sum(acc_r); acc_g = warp_sum(acc_g); acc_b = warp_sum(acc_b);
    acc_d = warp_sum(acc_d); acc_w = warp_sum(acc_w);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) {
      if (white_bkgd) { const float bg = 1.f - acc_w; acc_r += bg; acc_g += bg; acc_b += bg; }
      rgb[3 * (size_t)r + 0] = acc_r; rgb[3 * (size_t)r + 1] = acc_g; rgb[3 * (size_t)r + 2] = acc_b;
      depth[r] = acc_d;
      ray_mask[r] = cnt > 8 ? 1 : 0;     // render_ray.py:159
    }
  }
}

// Backward.  With f_s = 1 - a_s + 1e-10, T_s = prod_{j<s} f_j, w_s = a_s T_s and
// G_s = dL/dw_s (from rgb, depth, weights, white background):
//   dL/da_s = G_s T_s + d_alpha_s - (sum_{j>s} G_j w_j) / f_s ;   dL/dsigma_s = dL/da_s * exp(-sigma_s).
__global__ void __launch_bounds__(CMP_WARPS * 32)
k_composite_bwd(int R, int S, int white_bkgd, const float* __restrict__ raw, const float* __restrict__ z,
                const float* __restrict__ d_rgb, const float* __restrict__ d_depth,
                const float* __restrict__ d_weights, const float* __restrict__ d_alpha,
                float* __restrict__ d_raw) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* sT = smem + (size_t)wid * 2 * S;   // transmittance
  float* sG = sT + S;                       // G_s * w_s
  const int wglobal = blockIdx.x * CMP_WARPS + wid;
  const int wstride = gridDim.x * CMP_WARPS;
  for (int r = wglobal; r < R; r += wstride) {
    float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f;
    if (d_rgb) { gr = __ldg(d_rgb + 3 * (size_t)r); gg = __ldg(d_rgb + 3 * (size_t)r + 1); gb = __ldg(d_rgb + 3 * (size_t)r + 2); }
    if (d_depth) gd = __ldg(d_depth + r);
    const float bg = white_bkgd ? -(gr + gg + gb) : 0.f;
    double carry = 1.0;
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      const bool in = s < S;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      float zz = 0.f;
      if (in) { q = __ldg(reinterpret_cast<const float4*>(raw) + (size_t)r * S + s); zz = __ldg(z + (size_t)r * S + s); }
      const float a = in ? (1.f - expf(-q.w)) : 0.f;
      const float f = in ? __fadd_rn(__fsub_rn(1.f, a), 1e-10f) : 1.f;
      const double incl = warp_incl_prod((double)f, lane) * carry;
      double excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = carry;
      carry = __shfl_sync(0xffffffffu, incl, 31);
      if (in) {
        const float T = (float)excl;
        float G = gr * q.x + gg * q.y + gb * q.z + gd * zz + bg;
        if (d_weights) G += __ldg(d_weights + (size_t)r * S + s);
        sT[s] = T;
        sG[s] = G * a * T;
      }
    }
    __syncwarp();
    float tail = 0.f;                        // sum of G_j w_j over all later chunks
    for (int s0 = ((S - 1) / 32) * 32; s0 >= 0; s0 -= 32) {
      const int s = s0 + lane;
      const bool in = s < S;
      const float gw = in ? sG[s] : 0.f;
      float tot;
      const float suf = warp_suffix_excl_sum(gw, lane, tot) + tail;
      tail += tot;
      if (in) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(raw) + (size_t)r * S + s);
        const float e = expf(-q.w);
        const float a = 1.f - e;
        const float f = __fadd_rn(__fsub_rn(1.f, a), 1e-10f);
        const float T = sT[s];
        const float w = a * T;
        const float zz = __ldg(z + (size_t)r * S + s);
        float Gs = gr * q.x + gg * q.y + gb * q.z + gd * zz + bg;
        if (d_weights) Gs += __ldg(d_weights + (size_t)r * S + s);
        float da = Gs * T - suf / f;
        if (d_alpha) da += __ldg(d_alpha + (size_t)r * S + s);
        reinterpret_cast<float4*>(d_raw)[(size_t)r * S + s] = make_float4(w * gr, w * gg, w * gb, da * e);
      }
    }
    __syncwarp();
  }
}

extern "C" int nfb_composite_fwd(int R, int S, int white_bkgd, const float* raw, const float* z,
                                 const uint8_t* pixel_mask, const float* n_valid, int n_valid_stride,
                                 float* rgb, float* depth, float* weights, float* alpha, uint8_t* ray_mask,
                                 void* stream) {
  NFB_REQUIRE(R >= 0 && S >= 1, NFB_EINVAL, "nfb_composite_fwd: bad arguments (R=%d S=%d)", R, S);
  if (R == 0) return NFB_OK;
  NFB_REQUIRE(raw && z && rgb && depth && weights && alpha && ray_mask, NFB_EINVAL, "nfb_composite_fwd: NULL buffer");
  NFB_REQUIRE(pixel_mask || n_valid, NFB_EINVAL, "nfb_composite_fwd: need pixel_mask or n_valid");
  NFB_REQUIRE(((uintptr_t)raw % 16) == 0, NFB_EINVAL, "nfb_composite_fwd: raw must be 16-byte aligned");
  if (R == 0) return NFB_OK;
  NFB_RESOLVE_ONCE(k_composite_fwd, "nfb_composite_fwd");
  const int grid = tiles_grid((size_t)R, CMP_WARPS, 16);
  k_composite_fwd<<<grid, CMP_WARPS * 32, 0, (cudaStream_t)stream>>>(R, S, white_bkgd, raw, z, pixel_mask, n_valid,
                                                                  n_valid_stride, rgb, depth, weights, alpha, ray_mask);
  NFB_CHECK_LAUNCH("k_composite_fwd");
  return NFB_OK;
}

extern "C" int nfb_composite_bwd(int R, int S, int white_bkgd, const float* raw, const float* z,
                                 const float* d_rgb, const float* d_depth, const float* d_weights,
                                 const float* d_alpha, float* d_raw, void* stream) {
  NFB_REQUIRE(R >= 0 && S >= 1, NFB_EINVAL, "nfb_composite_bwd: bad arguments (R=%d S=%d)", R, S);
  if (R == 0) return NFB_OK;
  NFB_REQUIRE(raw && z && d_raw, NFB_EINVAL, "nfb_composite_bwd: NULL buffer");
  NFB_REQUIRE(((uintptr_t)raw % 16) == 0 && ((uintptr_t)d_raw % 16) == 0, NFB_EINVAL,
              "nfb_composite_bwd: raw / d_raw must be 16-byte aligned");
  NFB_REQUIRE(S <= 4096, NFB_EUNSUPPORTED, "nfb_composite_bwd: S=%d > 4096", S);
  if (R == 0) return NFB_OK;
  const int grid = tiles_grid((size_t)R, CMP_WARPS, 16);
  const size_t smem = (size_t)CMP_WARPS * 2 * S * sizeof(float);   // 32 S bytes: S = 4096 -> 128 KB (limit 227 KB)
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_composite_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return nfb_set_error(NFB_ECUDA, "nfb_composite_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  k_composite_bwd<<<grid, CMP_WARPS * 32, smem, (cudaStream_t)stream>>>(R, S, white_bkgd, raw, z, d_rgb, d_depth,
                                                                     d_weights, d_alpha, d_raw);
  NFB_CHECK_LAUNCH("k_composite_bwd");
  return NFB_OK;
}

// =====================================================================================================
// sample_pdf (render_ray.py:24-70) and the fine-depth construction around it (render_ray.py:216-238).
// One warp per ray.  The normaliser and the running CDF are accumulated in fp64 and rounded to fp32 per
// entry (torch's CPU cumsum does exactly that; see oracle/ibrnet_oracle.py for the normaliser).
// =====================================================================================================
// Build cdf[0..M] in shared memory from weights w[0..M) (w_i read through `getw`).
template <class GetW>
__device__ __forceinline__ void build_cdf(int M, GetW getw, float* s_cdf, int lane) {
  double part = 0.0;
  for (int i = lane; i < M; i += 32) part += (double)__fadd_rn(getw(i), 1e-5f);
  const float total = (float)warp_sum_d(part);
  double carry = 0.0;
  for (int i0 = 0; i0 < M; i0 += 32) {
    const int i = i0 + lane;
    const double pdf = (i < M) ? (double)__fdiv_rn(__fadd_rn(getw(i), 1e-5f), total) : 0.0;
    double incl = pdf;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    incl += carry;
    if (i < M) s_cdf[i + 1] = (float)incl;
    carry = __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) s_cdf[0] = 0.f;
  __syncwarp();
}

// Invert the CDF for one u (render_ray.py:47-68).  `above` counts cdf[0..M) entries <= u.
__device__ __forceinline__ float invert_one(float u, int M, const float* s_cdf, const float* s_bins, int& above_out) {
  int lo = 0, hi = M;               // cdf is non-decreasing: count = first index in [0,M) with cdf > u
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (u >= s_cdf[mid]) lo = mid + 1; else hi = mid;
  }
  const int above = lo;
  const int below = above - 1 > 0 ? above - 1 : 0;
  const float c0 = s_cdf[below], c1 = s_cdf[above];
  const float b0 = s_bins[below], b1 = s_bins[above];
  float denom = __fsub_rn(c1, c0);
  if (denom < 1e-5f) denom = 1.f;
  const float t = __fdiv_rn(__fsub_rn(u, c0), denom);
  above_out = above;
  return __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
}

constexpr int SP_WARPS = 4;

__global__ void __launch_bounds__(SP_WARPS * 32)
k_sample_pdf(int R, int M, int n, const float* __restrict__ bins, const float* __restrict__ weights,
             const float* __restrict__ u, int u_rows, float* __restrict__ samples, int64_t* __restrict__ above) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* s_cdf = smem + (size_t)wid * 2 * (M + 1);
  float* s_bins = s_cdf + (M + 1);
  for (int r = blockIdx.x * SP_WARPS + wid; r < R; r += gridDim.x * SP_WARPS) {
    const float* wr = weights + (size_t)r * M;
    build_cdf(M, [&](int i) { return __ldg(wr + i); }, s_cdf, lane);
    for (int i = lane; i <= M; i += 32) s_bins[i] = __ldg(bins + (size_t)r * (M + 1) + i);
    __syncwarp();
    const float* ur = u + (u_rows == 1 ? 0 : (size_t)r * n);
    for (int j = lane; j < n; j += 32) {
      int ab;
      samples[(size_t)r * n + j] = invert_one(__ldg(ur + j), M, s_cdf, s_bins, ab);
      if (above) above[(size_t)r * n + j] = ab;
    }
    __syncwarp();
  }
}

// Fine depths: bins = mid-points of z (or flipped mid-points of 1/z), weights = coarse weights[1:-1]
// (flipped with the bins), new samples (inverted back if inv_uniform), concatenated with the coarse
// depths and sorted ascending (rank sort: values only matter, ties broken by position).
__global__ void __launch_bounds__(SP_WARPS * 32)
k_fine_depths(int R, int S, int n_imp, int inv_uniform, const float* __restrict__ z_coarse,
              const float* __restrict__ w_coarse, const float* __restrict__ u, int u_rows,
              float* __restrict__ z_fine) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int M = S - 2, T = S + n_imp;
  float* s_cdf = smem + (size_t)wid * (2 * (M + 1) + T);
  float* s_bins = s_cdf + (M + 1);
  float* s_all = s_bins + (M + 1);
  for (int r = blockIdx.x * SP_WARPS + wid; r < R; r += gridDim.x * SP_WARPS) {
    const float* zr = z_coarse + (size_t)r * S;
    const float* wr = w_coarse + (size_t)r * S;
    // bins (M+1 = S-1 entries)
    for (int i = lane; i < S - 1; i += 32) {
      if (inv_uniform) {
        const int k = S - 2 - i;   // flipped index into the mid-points of 1/z
        s_bins[i] = __fmul_rn(0.5f, __fadd_rn(__fdiv_rn(1.f, __ldg(zr + k + 1)), __fdiv_rn(1.f, __ldg(zr + k))));
      } else {
        s_bins[i] = __fmul_rn(0.5f, __fadd_rn(__ldg(zr + i + 1), __ldg(zr + i)));
      }
    }
    build_cdf(M, [&](int i) { return inv_uniform ? __ldg(wr + 1 + (M - 1 - i)) : __ldg(wr + 1 + i); }, s_cdf, lane);
    for (int i = lane; i < S; i += 32) s_all[i] = __ldg(zr + i);
    const float* ur = u + (u_rows == 1 ? 0 : (size_t)r * n_imp);
    for (int j = lane; j < n_imp; j += 32) {
      int ab;
      float smp = invert_one(__ldg(ur + j), M, s_cdf, s_bins, ab);
      if (inv_uniform) smp = __fdiv_rn(1.f, smp);
      s_all[S + j] = smp;
    }
    __syncwarp();
    // torch.sort of cat(z_coarse, z_new) (render_ray.py:235-238), stable on ties (lower index first).  The coarse depths
    // are non-decreasing and, for the deterministic sampler, the new samples are monotone (increasing u -> increasing
    // inverse depth -> decreasing z with inv_uniform), so the sort is a merge of two monotone runs: the rank of an
    // element is its position in its own run plus a binary search in the other one.  Checked per ray with a warp vote;
    // anything else (stochastic u) takes the O(T^2) rank sort.
    const float* A = s_all;            // [S]
    const float* B = s_all + S;        // [n_imp]
    bool okA = true, inc = true, dec = true;
    for (int i = lane; i + 1 < S; i += 32) okA = okA && (A[i] <= A[i + 1]);
    for (int j = lane; j + 1 < n_imp; j += 32) { inc = inc && (B[j] <= B[j + 1]); dec = dec && (B[j] >= B[j + 1]); }
    okA = __all_sync(0xffffffffu, okA);
    inc = __all_sync(0xffffffffu, inc);
    dec = __all_sync(0xffffffffu, dec);
    if (okA && (inc || dec)) {
      for (int i = lane; i < T; i += 32) {
        const float v = s_all[i];
        int rank;
        if (i < S) {
          // coarse element: i earlier coarse elements + the new samples strictly below it (they lose ties)
          int lo = 0, hi = n_imp;      // count of B < v
          if (inc) { while (lo < hi) { const int mid = (lo + hi) >> 1; if (B[mid] < v) lo = mid + 1; else hi = mid; } rank = i + lo; }
          else     { while (lo < hi) { const int mid = (lo + hi) >> 1; if (B[mid] >= v) lo = mid + 1; else hi = mid; } rank = i + (n_imp - lo); }
        } else {
          const int j = i - S;
          int lo = 0, hi = S;          // coarse elements <= v come first (they win ties)
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (A[mid] <= v) lo = mid + 1; else hi = mid; }
          if (inc) rank = lo + j;
          else {
            int l2 = 0, h2 = n_imp;    // #{B > v}: the equal run of v starts here
            while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (B[mid] > v) l2 = mid + 1; else h2 = mid; }
            int l3 = l2, h3 = n_imp;   // #{B >= v}
            while (l3 < h3) { const int mid = (l3 + h3) >> 1; if (B[mid] >= v) l3 = mid + 1; else h3 = mid; }
            rank = lo + (n_imp - l3) + (j - l2);
          }
        }
        z_fine[(size_t)r * T + rank] = v;
      }
    } else {
      for (int i = lane; i < T; i += 32) {
        const float vi = s_all[i];
        int rank = 0;
        for (int j = 0; j < T; ++j) {
          const float vj = s_all[j];
          rank += (vj < vi || (vj == vi && j < i)) ? 1 : 0;
        }
        z_fine[(size_t)r * T + rank] = vi;
      }
    }
    __syncwarp();
  }
}

extern "C" int nfb_sample_pdf(int R, int M, int n, const float* bins, const float* weights, const float* u,
                              int u_rows, float* samples, int64_t* above, void* stream) {
  NFB_REQUIRE(R >= 0 && M >= 1 && n >= 1, NFB_EINVAL, "nfb_sample_pdf: bad arguments (R=%d M=%d n=%d)", R, M, n);
  if (R == 0) return NFB_OK;
  NFB_REQUIRE(bins && weights && u && samples, NFB_EINVAL, "nfb_sample_pdf: NULL buffer");
  NFB_REQUIRE(u_rows == 1 || u_rows == R, NFB_EINVAL, "nfb_sample_pdf: u_rows must be 1 or R");
  NFB_REQUIRE(M <= 4096, NFB_EUNSUPPORTED, "nfb_sample_pdf: M=%d > 4096 bins", M);
  if (R == 0) return NFB_OK;
  const size_t smem = (size_t)SP_WARPS * 2 * (M + 1) * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(k_sample_pdf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_sample_pdf<<<tiles_grid((size_t)R, SP_WARPS, 16), SP_WARPS * 32, smem, (cudaStream_t)stream>>>(
      R, M, n, bins, weights, u, u_rows, samples, above);
  NFB_CHECK_LAUNCH("k_sample_pdf");
  return NFB_OK;
}

extern "C" int nfb_fine_depths(int R, int S, int n_imp, int inv_uniform, const float* z_coarse,
                               const float* weights_coarse, const float* u, int u_rows, float* z_fine,
                               void* stream) {
  NFB_REQUIRE(R >= 0 && S >= 3 && n_imp >= 1, NFB_EINVAL, "nfb_fine_depths: bad arguments (R=%d S=%d n_imp=%d)", R, S, n_imp);
  if (R == 0) return NFB_OK;
  NFB_REQUIRE(z_coarse && weights_coarse && u && z_fine, NFB_EINVAL, "nfb_fine_depths: NULL buffer");
  NFB_REQUIRE(u_rows == 1 || u_rows == R, NFB_EINVAL, "nfb_fine_depths: u_rows must be 1 or R");
  NFB_REQUIRE(S + n_imp <= 2048, NFB_EUNSUPPORTED, "nfb_fine_depths: S+n_imp=%d > 2048", S + n_imp);
  if (R == 0) return NFB_OK;
  const size_t smem = (size_t)SP_WARPS * (2 * (S - 1) + S + n_imp) * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(k_fine_depths, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_fine_depths<<<tiles_grid((size_t)R, SP_WARPS, 16), SP_WARPS * 32, smem, (cudaStream_t)stream>>>(
      R, S, n_imp, inv_uniform, z_coarse, weights_coarse, u, u_rows, z_fine);
  NFB_CHECK_LAUNCH("k_fine_depths");
  return NFB_OK;
}
