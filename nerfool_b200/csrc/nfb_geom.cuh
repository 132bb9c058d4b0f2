// Per-(point, view) geometry shared by the standalone gather kernels and the fused view stage:
// camera projection + in-frustum mask (projection.py:24-62), ray_diff (projection.py:64-87) and the
// bilinear tap set of F.grid_sample(bilinear, zeros, align_corners=True) (projection.py:112-123).
#pragma once
#include "nfb_common.cuh"

// Where the 3-D points of a launch come from: an explicit [N][3] tensor, or z*ray_d + ray_o.
struct PointSrc {
  const float* xyz;    // [N][3] or nullptr
  const float* ray_o;  // [R][3]
  const float* ray_d;  // [R][3]
  const float* z;      // [R][S]
  int S;
};

__device__ __forceinline__ void load_point(const PointSrc& ps, int p, float& x, float& y, float& z) {
  if (ps.xyz) {
    x = __ldg(ps.xyz + 3 * (size_t)p + 0);
    y = __ldg(ps.xyz + 3 * (size_t)p + 1);
    z = __ldg(ps.xyz + 3 * (size_t)p + 2);
  } else {
    // pts = z * ray_d + ray_o as two separately rounded ops (render_ray.py:115,241-243): no FMA contraction
    const int r = p / ps.S;
    const float t = __ldg(ps.z + p);
    x = __fadd_rn(__fmul_rn(t, __ldg(ps.ray_d + 3 * r + 0)), __ldg(ps.ray_o + 3 * r + 0));
    y = __fadd_rn(__fmul_rn(t, __ldg(ps.ray_d + 3 * r + 1)), __ldg(ps.ray_o + 3 * r + 1));
    z = __fadd_rn(__fmul_rn(t, __ldg(ps.ray_d + 3 * r + 2)), __ldg(ps.ray_o + 3 * r + 2));
  }
}

struct ViewGeom {
  float gx, gy;        // normalised grid coordinates in [-1,1] (projection.py:37-40)
  float mask;          // inbound & in-front, as 0/1 float (projection.py:128-131)
  float rd[4];         // ray_diff (projection.py:64-87)
};

// cam16: 16 floats of view v (P rows 0-2, centre); tgt: centre of the query camera.
__device__ __forceinline__ ViewGeom view_geometry(float x, float y, float z, const float* __restrict__ cam16,
                                                  const float* __restrict__ tgt, float Wm1, float Hm1) {
  ViewGeom g;
  // P*[x,y,z,1]: the CPU bmm evaluates each row as fma(p3,1, fma(p2,z, fma(p1,y, p0*x))) (measured 100 %)
  const float px = __fadd_rn(__fmaf_rn(cam16[2], z, __fmaf_rn(cam16[1], y, __fmul_rn(cam16[0], x))), cam16[3]);
  const float py = __fadd_rn(__fmaf_rn(cam16[6], z, __fmaf_rn(cam16[5], y, __fmul_rn(cam16[4], x))), cam16[7]);
  const float pz = __fadd_rn(__fmaf_rn(cam16[10], z, __fmaf_rn(cam16[9], y, __fmul_rn(cam16[8], x))), cam16[11]);
  const float den = fmaxf(pz, 1e-8f);
  float u = __fdiv_rn(px, den);
  float v = __fdiv_rn(py, den);
  u = fminf(fmaxf(u, -1e6f), 1e6f);
  v = fminf(fmaxf(v, -1e6f), 1e6f);
  const bool ok = (u <= Wm1) && (u >= 0.f) && (v <= Hm1) && (v >= 0.f) && (pz > 0.f);
  g.mask = ok ? 1.f : 0.f;
  g.gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, u), Wm1), 1.f);
  g.gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, v), Hm1), 1.f);

  // ray_diff: a = unit(target centre - x), b = unit(source centre - x), each / (norm + 1e-6).
  // Rounding order reproduces torch's CPU kernels bit for bit (probed: norm over 3 = sqrt(fma(z,z,fma(y,y,x*x))),
  // sum(a*b) = (p0 + p1) + p2 with separately rounded products); the dot feeds an ill-conditioned
  // difference of exponentials in IBRNet's pooling weights, so ulp-level agreement matters here.
  float ax = __fsub_rn(tgt[0], x), ay = __fsub_rn(tgt[1], y), az = __fsub_rn(tgt[2], z);
  float bx = __fsub_rn(cam16[12], x), by = __fsub_rn(cam16[13], y), bz = __fsub_rn(cam16[14], z);
  const float an = __fadd_rn(__fsqrt_rn(__fmaf_rn(az, az, __fmaf_rn(ay, ay, __fmul_rn(ax, ax)))), 1e-6f);
  const float bn = __fadd_rn(__fsqrt_rn(__fmaf_rn(bz, bz, __fmaf_rn(by, by, __fmul_rn(bx, bx)))), 1e-6f);
  ax = __fdiv_rn(ax, an); ay = __fdiv_rn(ay, an); az = __fdiv_rn(az, an);
  bx = __fdiv_rn(bx, bn); by = __fdiv_rn(by, bn); bz = __fdiv_rn(bz, bn);
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  const float dn = fmaxf(__fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)))), 1e-6f);
  g.rd[0] = __fdiv_rn(dx, dn);
  g.rd[1] = __fdiv_rn(dy, dn);
  g.rd[2] = __fdiv_rn(dz, dn);
  g.rd[3] = __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
  return g;
}

// Bilinear tap set on a w x h source for normalised coords (gx, gy), align_corners=True, zero padding.
// off[i] is the texel index y*w+x of tap i (nw, ne, sw, se) or -1 if that tap is outside; wt[i] its weight.
// Weight construction follows ATen's CPU kernel: t = x - floor(x), e = 1 - t (GridSamplerKernel.cpp).
struct Taps {
  int off[4];
  float wt[4];
};
__device__ __forceinline__ Taps bilinear_taps(float gx, float gy, int w, int h) {
  Taps t;
  const float ix = __fmul_rn(__fadd_rn(gx, 1.f), 0.5f * (float)(w - 1));
  const float iy = __fmul_rn(__fadd_rn(gy, 1.f), 0.5f * (float)(h - 1));
  const float fx = floorf(ix), fy = floorf(iy);
  const float tw = ix - fx, te = 1.f - tw;   // distance to west / east
  const float tn = iy - fy, ts = 1.f - tn;   // distance to north / south
  t.wt[0] = ts * te; t.wt[1] = ts * tw; t.wt[2] = tn * te; t.wt[3] = tn * tw;
  // |ix| can be up to ~1e6*w after the clamp in view_geometry: saturating float->int conversion is safe
  const int x0 = (int)fx, y0 = (int)fy;
  const int x1 = x0 + 1, y1 = y0 + 1;
  const bool xw = (x0 >= 0) && (x0 < w), xe = (x1 >= 0) && (x1 < w);
  const bool yn = (y0 >= 0) && (y0 < h), ys = (y1 >= 0) && (y1 < h);
  t.off[0] = (xw && yn) ? y0 * w + x0 : -1;
  t.off[1] = (xe && yn) ? y0 * w + x1 : -1;
  t.off[2] = (xw && ys) ? y1 * w + x0 : -1;
  t.off[3] = (xe && ys) ? y1 * w + x1 : -1;
  return t;
}

// Gather the 35-channel row of one (point, view): 3 RGB from imgs[V][H][W][3] + 32 features from the
// channel-last map feat[V][fh][fw][32] (each tap = one 128-byte line = 8 x LDG.128).
__device__ __forceinline__ void gather_row(const ViewGeom& g, int v, int H, int W, int fh, int fw,
                                           const float* __restrict__ imgs, const float* __restrict__ feat,
                                           float (&row)[NFB_ROW_CH]) {
#pragma unroll
  for (int c = 0; c < NFB_ROW_CH; ++c) row[c] = 0.f;
  {
    const Taps t = bilinear_taps(g.gx, g.gy, W, H);
    const float* base = imgs + (size_t)v * H * W * 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (t.off[i] >= 0) {
        const float* p = base + (size_t)t.off[i] * 3;
        row[0] += __ldg(p + 0) * t.wt[i];
        row[1] += __ldg(p + 1) * t.wt[i];
        row[2] += __ldg(p + 2) * t.wt[i];
      }
    }
  }
  {
    const Taps t = bilinear_taps(g.gx, g.gy, fw, fh);
    const float4* base = reinterpret_cast<const float4*>(feat + (size_t)v * fh * fw * NFB_FEAT_CH);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (t.off[i] >= 0) {
        const float4* p = base + (size_t)t.off[i] * (NFB_FEAT_CH / 4);
        const float wgt = t.wt[i];
#pragma unroll
        for (int j = 0; j < NFB_FEAT_CH / 4; ++j) {
          const float4 q = __ldg(p + j);
          // packed fp32x2: the same four FMAs (bit-identical), half the issue slots
          const float2 w2 = make_float2(wgt, wgt);
          const float2 r0 = __ffma2_rn(make_float2(q.x, q.y), w2, make_float2(row[3 + 4 * j + 0], row[3 + 4 * j + 1]));
          const float2 r1 = __ffma2_rn(make_float2(q.z, q.w), w2, make_float2(row[3 + 4 * j + 2], row[3 + 4 * j + 3]));
          row[3 + 4 * j + 0] = r0.x; row[3 + 4 * j + 1] = r0.y; row[3 + 4 * j + 2] = r1.x; row[3 + 4 * j + 3] = r1.y;
        }
      }
    }
  }
}

// Warp-cooperative form of gather_row for kernels whose 32 lanes hold 32 different rows (nfb_view_tc.cuh).  A lane's own
// LDG.128 covers 16 of the 128 bytes of a feature line, so a warp-wide load touches ~30 different lines for 512 useful bytes and
// the same lines are revisited by the seven other loads of the tap: the L1 data pipe is the most loaded unit of the kernel
// (l1tex__data_pipe_lsu_wavefronts 64 %) and the consumer of these loads holds the largest share of its stall samples.
// Here the 8 lanes of a quarter-warp read one row's four feature lines TOGETHER (lane k the channels 4k..4k+3 of every tap): a
// warp-wide load is 4 whole lines.  Tap offsets / weights of the row come from its owner lane by shuffle; after 8 rounds every
// lane holds the channel quad k of the 8 rows of its quarter-warp, which are handed to the owners through `stage`
// (row r: floats [4, 36), 16-byte aligned; any per-row scratch with that room, row stride `stride` floats).  Same taps, weights and
// summation order as gather_row: bit-identical rows.  Every lane of the warp must call (inactive rows: active = false).
__device__ __forceinline__ void gather_row_coop(bool active, const ViewGeom& g, int v, int H, int W, int fh, int fw,
                                                const float* __restrict__ imgs, const float* __restrict__ feat,
                                                float* __restrict__ stage, int stride, int my_row, float (&row)[NFB_ROW_CH]) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, k = lane & 7, q0 = lane & ~7;
  row[0] = row[1] = row[2] = 0.f;
  if (active) {
    const Taps t = bilinear_taps(g.gx, g.gy, W, H);
    const float* base = imgs + (size_t)v * H * W * 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (t.off[i] >= 0) {
        const float* p = base + (size_t)t.off[i] * 3;
        row[0] += __ldg(p + 0) * t.wt[i];
        row[1] += __ldg(p + 1) * t.wt[i];
        row[2] += __ldg(p + 2) * t.wt[i];
      }
    }
  }
  Taps t = bilinear_taps(active ? g.gx : 0.f, active ? g.gy : 0.f, fw, fh);
  if (!active) { t.off[0] = t.off[1] = t.off[2] = t.off[3] = -1; }
  // texel index inside the whole [V][fh][fw] map, so one shuffled integer addresses the line
  const int plane = fh * fw;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (t.off[i] >= 0) t.off[i] += v * plane;
  const float4* fbase = reinterpret_cast<const float4*>(feat) + k;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int src = q0 + it;
    float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int off = __shfl_sync(FULL, t.off[i], src);
      const float wgt = __shfl_sync(FULL, t.wt[i], src);
      if (off >= 0) {
        const float4 qv = __ldg(fbase + (size_t)off * (NFB_FEAT_CH / 4));
        const float2 w2 = make_float2(wgt, wgt);
        a01 = __ffma2_rn(make_float2(qv.x, qv.y), w2, a01);
        a23 = __ffma2_rn(make_float2(qv.z, qv.w), w2, a23);
      }
    }
    *reinterpret_cast<float4*>(stage + (size_t)(my_row - lane + src) * stride + 4 + 4 * k) = make_float4(a01.x, a01.y, a23.x, a23.y);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < NFB_FEAT_CH / 4; ++j) {
    const float4 qv = *reinterpret_cast<const float4*>(stage + (size_t)my_row * stride + 4 + 4 * j);
    row[3 + 4 * j + 0] = qv.x; row[3 + 4 * j + 1] = qv.y; row[3 + 4 * j + 2] = qv.z; row[3 + 4 * j + 3] = qv.w;
  }
  __syncwarp();                  // the staging rows are reused by the caller
}

// Scatter-add the cotangent of one gathered row back into d_feat / d_imgs (grid_sampler_2d backward
// w.r.t. the input).  Each feature tap is 8 x RED.128 (vector float atomics, sm_90+).
__device__ __forceinline__ void scatter_row(const ViewGeom& g, int v, int H, int W, int fh, int fw,
                                            const float (&d_row)[NFB_ROW_CH], float* __restrict__ d_feat,
                                            float* __restrict__ d_imgs) {
  if (d_imgs) {
    const Taps t = bilinear_taps(g.gx, g.gy, W, H);
    float* base = d_imgs + (size_t)v * H * W * 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (t.off[i] >= 0) {
        float* p = base + (size_t)t.off[i] * 3;
        atomicAdd(p + 0, d_row[0] * t.wt[i]);
        atomicAdd(p + 1, d_row[1] * t.wt[i]);
        atomicAdd(p + 2, d_row[2] * t.wt[i]);
      }
    }
  }
  if (d_feat) {
    const Taps t = bilinear_taps(g.gx, g.gy, fw, fh);
    float4* base = reinterpret_cast<float4*>(d_feat + (size_t)v * fh * fw * NFB_FEAT_CH);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (t.off[i] >= 0) {
        float4* p = base + (size_t)t.off[i] * (NFB_FEAT_CH / 4);
        const float wgt = t.wt[i];
#pragma unroll
        for (int j = 0; j < NFB_FEAT_CH / 4; ++j) {
          float4 q;
          q.x = d_row[3 + 4 * j + 0] * wgt;
          q.y = d_row[3 + 4 * j + 1] * wgt;
          q.z = d_row[3 + 4 * j + 2] * wgt;
          q.w = d_row[3 + 4 * j + 3] * wgt;
          atomicAdd(p + j, q);
        }
      }
    }
  }
}

// Warp-cooperative scatter, the mirror of gather_row_coop: the 8 lanes of a quarter-warp add one row's weighted cotangent to the
// row's four feature lines TOGETHER (lane k the channels 4k..4k+3 of every tap), so a warp-wide RED.128 covers 4 whole 128-byte
// lines of the gradient map instead of 16-byte pieces of ~30 different lines.  The cotangent rows are handed from their owners
// to the quarter-warp through `stage` (row r: floats [4, 36)); tap offsets / weights by shuffle.  Every lane of the warp must call.
__device__ __forceinline__ void scatter_row_coop(bool active, float gx, float gy, int v, int H, int W, int fh, int fw,
                                                 const float (&d_row)[NFB_ROW_CH], float* __restrict__ d_feat, float* __restrict__ d_imgs,
                                                 float* __restrict__ stage, int stride, int my_row) {
  if (active && d_imgs) {
    const Taps t = bilinear_taps(gx, gy, W, H);
    float* base = d_imgs + (size_t)v * H * W * 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (t.off[i] >= 0) {
        float* p = base + (size_t)t.off[i] * 3;
        atomicAdd(p + 0, d_row[0] * t.wt[i]);
        atomicAdd(p + 1, d_row[1] * t.wt[i]);
        atomicAdd(p + 2, d_row[2] * t.wt[i]);
      }
    }
  }
  if (!d_feat) return;                                   // kernel argument: uniform
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, k = lane & 7, q0 = lane & ~7;
#pragma unroll
  for (int j = 0; j < NFB_FEAT_CH / 4; ++j)
    *reinterpret_cast<float4*>(stage + (size_t)my_row * stride + 4 + 4 * j) =
        make_float4(d_row[3 + 4 * j], d_row[4 + 4 * j], d_row[5 + 4 * j], d_row[6 + 4 * j]);
  Taps t = bilinear_taps(active ? gx : 0.f, active ? gy : 0.f, fw, fh);
  if (!active) { t.off[0] = t.off[1] = t.off[2] = t.off[3] = -1; }
  const int plane = fh * fw;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (t.off[i] >= 0) t.off[i] += v * plane;
  __syncwarp();
  float4* fbase = reinterpret_cast<float4*>(d_feat) + k;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int src = q0 + it;
    const float4 d4 = *reinterpret_cast<const float4*>(stage + (size_t)(my_row - lane + src) * stride + 4 + 4 * k);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int off = __shfl_sync(FULL, t.off[i], src);
      const float wgt = __shfl_sync(FULL, t.wt[i], src);
      if (off >= 0) atomicAdd(fbase + (size_t)off * (NFB_FEAT_CH / 4), make_float4(d4.x * wgt, d4.y * wgt, d4.z * wgt, d4.w * wgt));
    }
  }
  __syncwarp();
}

// Scatter with pairwise de-duplication.  Consecutive samples of a ray fall into the same texel quad of a source view most
// of the time (measured on the headline scene: 74 % at the coarse level, 87 % at the fine level).  The rows of a sample
// pair (2k, 2k + 1) of one view are the lanes (l, l ^ V) of a warp when V is a power of two <= 16: if both hit the same
// four texels, the even lane adds both weighted cotangents with ONE set of 8 x 4 RED.128 and the odd lane issues none --
// ~40 % fewer vector atomics into the L2-resident gradient map for 41 shuffles per row.
// Every lane of the warp must call (inactive rows pass active = false).  `partner` = lane of the same view of the paired
// sample (own lane when the row has no partner), `leader` = true for the lane that emits for a merged pair.  (For V a power
// of two that divides 32 the pairs are the lanes (l, l ^ V); the packed row mapping of nfb_view_tc.cuh pairs the samples
// (2k, 2k + 1) of a warp for any V <= 16.)
__device__ __forceinline__ void scatter_row_paired(bool active, float gx, float gy, int v, int partner, bool leader, int H, int W, int fh, int fw,
                                                   const float (&d_row)[NFB_ROW_CH], float* __restrict__ d_feat,
                                                   float* __restrict__ d_imgs) {
  if (active && d_imgs) {
    const Taps t = bilinear_taps(gx, gy, W, H);
    float* base = d_imgs + (size_t)v * H * W * 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (t.off[i] >= 0) {
        float* p = base + (size_t)t.off[i] * 3;
        atomicAdd(p + 0, d_row[0] * t.wt[i]);
        atomicAdd(p + 1, d_row[1] * t.wt[i]);
        atomicAdd(p + 2, d_row[2] * t.wt[i]);
      }
    }
  }
  if (!d_feat) return;                                   // kernel argument: uniform
  Taps t = bilinear_taps(active ? gx : 0.f, active ? gy : 0.f, fw, fh);
  if (!active) { t.off[0] = t.off[1] = t.off[2] = t.off[3] = -1; }
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int act_nb = __shfl_sync(FULL, (int)active, partner);
  bool same = active && (act_nb != 0) && (partner != lane);
  float wn[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int off_nb = __shfl_sync(FULL, t.off[i], partner);    // unconditionally: every lane must execute every shuffle
    same = same && (off_nb == t.off[i]);
    wn[i] = __shfl_sync(FULL, t.wt[i], partner);
  }
  const bool emit = active && !(same && !leader);
  float4* base = reinterpret_cast<float4*>(d_feat + (size_t)v * fh * fw * NFB_FEAT_CH);
#pragma unroll
  for (int j = 0; j < NFB_FEAT_CH / 4; ++j) {
    const float m0 = d_row[3 + 4 * j], m1 = d_row[4 + 4 * j], m2 = d_row[5 + 4 * j], m3 = d_row[6 + 4 * j];
    float n0 = __shfl_sync(FULL, m0, partner), n1 = __shfl_sync(FULL, m1, partner);
    float n2 = __shfl_sync(FULL, m2, partner), n3 = __shfl_sync(FULL, m3, partner);
    if (!same) { n0 = 0.f; n1 = 0.f; n2 = 0.f; n3 = 0.f; }       // the partner may be an inactive row holding anything
    if (emit) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (t.off[i] >= 0) {
          const float wm = t.wt[i], wp = same ? wn[i] : 0.f;
          const float2 wm2 = make_float2(wm, wm), wp2 = make_float2(wp, wp);
          const float2 q01 = __ffma2_rn(make_float2(n0, n1), wp2, __fmul2_rn(make_float2(m0, m1), wm2));
          const float2 q23 = __ffma2_rn(make_float2(n2, n3), wp2, __fmul2_rn(make_float2(m2, m3), wm2));
          const float4 q = make_float4(q01.x, q01.y, q23.x, q23.y);
          atomicAdd(base + (size_t)t.off[i] * (NFB_FEAT_CH / 4) + j, q);
        }
      }
    }
  }
}

