// Per-pixel math of the photometric chain (host+device so tests/hostcheck can evaluate it).
// Reference: depth_pose_prediction.py:976-1017 (_reconstruct_images), :1178-1192
// (_compute_reprojection_loss), networks/layers.py:51-137, utils.py:120-142.
#pragma once
#include "common.cuh"

namespace clslam {

constexpr int kPhotoTW = 32;   // tile width  (pixels)
constexpr int kPhotoTH = 16;   // tile height (pixels)
constexpr float kSsimC1 = 0.0001f;   // 0.01 ** 2
constexpr float kSsimC2 = 0.0009f;   // 0.03 ** 2

// ---- bilinear up-sampling, align_corners=False (F.interpolate, :989-991) -------------------
struct UpCoef { int i0, i1; float l0, l1; };
CLSLAM_HD UpCoef up_coef(int dst, int in_size, int out_size) {
  const float scale = (float)in_size / (float)out_size;
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  UpCoef c;
  c.i0 = (int)src;
  if (c.i0 > in_size - 1) c.i0 = in_size - 1;
  c.i1 = c.i0 + (c.i0 < in_size - 1 ? 1 : 0);
  float l1 = src - (float)c.i0;
  l1 = l1 < 0.f ? 0.f : (l1 > 1.f ? 1.f : l1);
  c.l1 = l1; c.l0 = 1.f - l1;
  return c;
}
CLSLAM_HD float disp_upsample(const float* disp_b, int hs, int ws, int h, int w, int y, int x) {
  const UpCoef cy = up_coef(y, hs, h), cx = up_coef(x, ws, w);
  const float* r0 = disp_b + (size_t)cy.i0 * ws;
  const float* r1 = disp_b + (size_t)cy.i1 * ws;
  return cy.l0 * (cx.l0 * ld1(r0 + cx.i0) + cx.l1 * ld1(r0 + cx.i1)) +
         cy.l1 * (cx.l0 * ld1(r1 + cx.i0) + cx.l1 * ld1(r1 + cx.i1));
}

CLSLAM_HD float depth_from_disp(float d, int mode, float min_depth, float max_depth) {
  if (mode == CLSLAM_DEPTH_INV) return 1.f / d;
  if (mode == CLSLAM_DEPTH_MIN_ONLY) return min_depth / d;
  const float lo = 1.f / max_depth, hi = 1.f / min_depth;
  return 1.f / (lo + (hi - lo) * d);
}
// d(depth)/d(disp)
CLSLAM_HD float depth_grad(float d, float depth, int mode, float min_depth, float max_depth) {
  if (mode == CLSLAM_DEPTH_MIN_MAX) {
    const float lo = 1.f / max_depth, hi = 1.f / min_depth;
    return -depth * depth * (hi - lo);
  }
  return -depth / d;
}

// ---- back-project, transform, project, un-normalise (layers.py:74-104 + grid_sample) -------
struct WarpGeom {
  float ray[3];     // inv_K[:3,:3] . (u, v, 1)
  float X[3];       // depth * ray
  float z;          // p_z + 1e-7
  float px, py;     // projected pixel coordinates p_xy / z
  float ix, iy;     // sampling coordinates after border clipping
  float mx, my;     // d(ix)/d(px), d(iy)/d(py): 0 where clipped (grid_sample backward), else 1
};
CLSLAM_HD void warp_geometry(const float* inv_k /*4x4*/, const float* P /*3x4*/, float depth, int u, int v, int h, int w,
                             WarpGeom& g) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    g.ray[r] = inv_k[r * 4 + 0] * (float)u + inv_k[r * 4 + 1] * (float)v + inv_k[r * 4 + 2];
    g.X[r] = depth * g.ray[r];
  }
  float p[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) p[r] = P[r * 4 + 0] * g.X[0] + P[r * 4 + 1] * g.X[1] + P[r * 4 + 2] * g.X[2] + P[r * 4 + 3];
  g.z = p[2] + 1e-7f;
  // one reciprocal instead of two divisions, and multiplications by 1/(W-1), 1/(H-1): <= 1 ulp from the
  // reference's division sequence (layers.py:98-103), far below the sampling-coordinate sensitivity
  const float rz = 1.f / g.z;
  g.px = p[0] * rz;
  g.py = p[1] * rz;
  const float gx = (g.px * (1.f / (float)(w - 1)) - 0.5f) * 2.f;
  const float gy = (g.py * (1.f / (float)(h - 1)) - 0.5f) * 2.f;
  float ix = ((gx + 1.f) * 0.5f) * (float)(w - 1);
  float iy = ((gy + 1.f) * 0.5f) * (float)(h - 1);
  // clip_coordinates_set_grad: gradient is zeroed where the coordinate is clipped
  g.mx = 1.f; g.my = 1.f;
  if (!(ix > 0.f)) { ix = 0.f; g.mx = 0.f; } else if (ix >= (float)(w - 1)) { ix = (float)(w - 1); g.mx = 0.f; }
  if (!(iy > 0.f)) { iy = 0.f; g.my = 0.f; } else if (iy >= (float)(h - 1)) { iy = (float)(h - 1); g.my = 0.f; }
  g.ix = ix; g.iy = iy;
}

// bilinear tap weights / indices of grid_sample(bilinear, border, align_corners=True)
struct Taps { int x0, y0; float nw, ne, sw, se; float tx0, tx1, ty0, ty1; bool x1ok, y1ok; };
CLSLAM_HD Taps make_taps(float ix, float iy, int h, int w) {
  Taps t;
  const float fx = floorf(ix), fy = floorf(iy);
  t.x0 = (int)fx; t.y0 = (int)fy;
  t.tx1 = ix - fx;            // ix - ix_nw
  t.tx0 = (fx + 1.f) - ix;    // ix_se - ix
  t.ty1 = iy - fy;
  t.ty0 = (fy + 1.f) - iy;
  t.nw = t.tx0 * t.ty0; t.ne = t.tx1 * t.ty0; t.sw = t.tx0 * t.ty1; t.se = t.tx1 * t.ty1;
  t.x1ok = t.x0 + 1 <= w - 1; t.y1ok = t.y0 + 1 <= h - 1;
  return t;
}
CLSLAM_HD float sample_plane(const float* plane, int w, const Taps& t) {
  const float* r0 = plane + (size_t)t.y0 * w + t.x0;
  float out = ld1(r0) * t.nw;
  if (t.x1ok) out += ld1(r0 + 1) * t.ne;
  if (t.y1ok) {
    out += ld1(r0 + w) * t.sw;
    if (t.x1ok) out += ld1(r0 + w + 1) * t.se;
  }
  return out;
}
// d(sample)/d(ix), d(sample)/d(iy) (before the clip mask)
CLSLAM_HD void sample_plane_grad(const float* plane, int w, const Taps& t, float& dix, float& diy) {
  const float* r0 = plane + (size_t)t.y0 * w + t.x0;
  const float nw = ld1(r0);
  const float ne = t.x1ok ? ld1(r0 + 1) : 0.f;
  const float sw = t.y1ok ? ld1(r0 + w) : 0.f;
  const float se = (t.x1ok && t.y1ok) ? ld1(r0 + w + 1) : 0.f;
  dix = -nw * t.ty0 + ne * t.ty0 - sw * t.ty1 + se * t.ty1;
  diy = -nw * t.tx0 - ne * t.tx1 + sw * t.tx0 + se * t.tx1;
}

// ---- the same sampling on RGBA-packed sources ((h, w, 4) floats): one 16-byte load per tap ----------
CLSLAM_HD void sample_rgba(const float* img, int w, const Taps& t, float (&out)[3]) {
  const float* r0 = img + (t.y0 * w + t.x0) * 4;
  const float4 nw = ld4(r0);
  out[0] = nw.x * t.nw; out[1] = nw.y * t.nw; out[2] = nw.z * t.nw;
  if (t.x1ok) { const float4 v = ld4(r0 + 4); out[0] += v.x * t.ne; out[1] += v.y * t.ne; out[2] += v.z * t.ne; }
  if (t.y1ok) {
    const float4 v = ld4(r0 + w * 4);
    out[0] += v.x * t.sw; out[1] += v.y * t.sw; out[2] += v.z * t.sw;
    if (t.x1ok) { const float4 u = ld4(r0 + w * 4 + 4); out[0] += u.x * t.se; out[1] += u.y * t.se; out[2] += u.z * t.se; }
  }
}
// sum_c g[c] * d(sample_c)/d(ix), .../d(iy) (before the clip mask)
CLSLAM_HD void sample_rgba_grad(const float* img, int w, const Taps& t, const float (&g)[3], float& gix, float& giy) {
  const float* r0 = img + (t.y0 * w + t.x0) * 4;
  const float4 nw = ld4(r0);
  const float4 ne = t.x1ok ? ld4(r0 + 4) : f4zero();
  const float4 sw = t.y1ok ? ld4(r0 + w * 4) : f4zero();
  const float4 se = (t.x1ok && t.y1ok) ? ld4(r0 + w * 4 + 4) : f4zero();
  const float dnw = g[0] * nw.x + g[1] * nw.y + g[2] * nw.z, dne = g[0] * ne.x + g[1] * ne.y + g[2] * ne.z;
  const float dsw = g[0] * sw.x + g[1] * sw.y + g[2] * sw.z, dse = g[0] * se.x + g[1] * se.y + g[2] * se.z;
  gix = -dnw * t.ty0 + dne * t.ty0 - dsw * t.ty1 + dse * t.ty1;
  giy = -dnw * t.tx0 - dne * t.tx1 + dsw * t.tx0 + dse * t.tx1;
}

// ---- SSIM of one channel from 3x3 window sums (networks/layers.py:123-137) -----------------
struct SsimStats { float mu_x, mu_y, a1, a2, b1, b2, n, d, val; };
CLSLAM_HD SsimStats ssim_stats(float sx, float sy, float sxx, float syy, float sxy) {
  SsimStats s;
  constexpr float k9 = 1.f / 9.f;      // AvgPool2d(3, 1) divides the window sum by 9
  s.mu_x = sx * k9; s.mu_y = sy * k9;
  const float sig_x = sxx * k9 - s.mu_x * s.mu_x;
  const float sig_y = syy * k9 - s.mu_y * s.mu_y;
  const float sig_xy = sxy * k9 - s.mu_x * s.mu_y;
  s.a1 = 2.f * s.mu_x * s.mu_y + kSsimC1;
  s.a2 = 2.f * sig_xy + kSsimC2;
  s.b1 = s.mu_x * s.mu_x + s.mu_y * s.mu_y + kSsimC1;
  s.b2 = sig_x + sig_y + kSsimC2;
  s.n = s.a1 * s.a2;
  s.d = s.b1 * s.b2;
  s.val = (1.f - s.n / s.d) * 0.5f;
  return s;
}
CLSLAM_HD float clamp01(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }

// 1-D multiplicity with which SSIM window centre p (in q-1..q+1) reads pixel q through the
// reflection padding (adjoint of ReflectionPad2d(1) + 3x3 box)
CLSLAM_HD float refl_box_weight(int q, int p, int size) {
  if (p < 0 || p >= size) return 0.f;
  return 1.f + ((q == 1 && p == 0) ? 1.f : 0.f) + ((q == size - 2 && p == size - 1) ? 1.f : 0.f);
}

}  // namespace clslam
