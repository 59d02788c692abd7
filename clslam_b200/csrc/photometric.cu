// Fused photometric chain: up-sample disparity -> depth -> back-project -> SE(3) -> project ->
// border-clipped bilinear warp -> SSIM + L1 -> auto-mask min, and its hand-derived backward.
// HBM-bound: each CTA owns a 32x16 pixel tile, stages the warped images of the tile + halo in
// shared memory once and evaluates all 3x3 stencils from there.  See include/clslam_b200.h.
#include <cstdlib>

#include "photo_ops.cuh"

namespace clslam {

constexpr int kPhotoThreads = 256;
// forward: 1-pixel halo
constexpr int kFW = kPhotoTW + 2, kFH = kPhotoTH + 2, kFN = kFW * kFH;
// backward: 2-pixel halo for the warped image, 1-pixel halo for the SSIM coefficient maps
constexpr int kBW = kPhotoTW + 4, kBH = kPhotoTH + 4, kBN = kBW * kBH;

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// image coordinate read by padded coordinate `i` (reflection, then clamped for tiles that overhang)
__device__ __forceinline__ int pad_coord(int i, int size) { return clampi(reflect_idx(i, size), 0, size - 1); }

__device__ __forceinline__ float block_sum(float v, float* red /*[8]*/) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kPhotoThreads / 32; ++i) s += red[i];
  }
  return s;   // valid on thread 0
}

// 0.85 * mean_c SSIM + 0.15 * mean_c L1 at local position (ly, lx) of the haloed tiles
template <int LW>
__device__ __forceinline__ float reproj_error(const float* xs /*[3][..][LW]*/, const float* ys, int plane, int ly, int lx) {
  float ssim_sum = 0.f, l1_sum = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* xp = xs + c * plane;
    const float* yp = ys + c * plane;
    float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const float xv = xp[(ly + dy) * LW + lx + dx], yv = yp[(ly + dy) * LW + lx + dx];
        sx += xv; sy += yv; sxx += xv * xv; syy += yv * yv; sxy += xv * yv;
      }
    ssim_sum += clamp01(ssim_stats(sx, sy, sxx, syy, sxy).val);
    l1_sum += fabsf(yp[ly * LW + lx] - xp[ly * LW + lx]);
  }
  return 0.85f * (ssim_sum / 3.f) + 0.15f * (l1_sum / 3.f);
}

// ---- (n, 3, h, w) planar -> (n, h, w, 4) packed: the warp gathers one float4 per bilinear tap ----------
__global__ void __launch_bounds__(256)
pack_rgba_kernel(const float* __restrict__ planar, float* __restrict__ packed, int n, int hw) {
  pdl_wait();
  pdl_trigger();
  const size_t total = (size_t)n * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t img = i / hw, px = i - img * hw;
    const float* p = planar + img * 3 * hw + px;
    *reinterpret_cast<float4*>(packed + i * 4) = make_float4(ld1(p), ld1(p + hw), ld1(p + 2 * (size_t)hw), 0.f);
  }
}

// ---- identity (auto-mask) candidates -------------------------------------------------------
__global__ void __launch_bounds__(kPhotoThreads)
photo_identity_kernel(const float* __restrict__ target, const float* __restrict__ sources, float* __restrict__ identity,
                      int b, int h, int w, int tiles_x) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tg[3 * kFN];
  __shared__ float sr[3 * kFN];
  const int bb = blockIdx.y;
  const int tx0 = (blockIdx.x % tiles_x) * kPhotoTW, ty0 = (blockIdx.x / tiles_x) * kPhotoTH;
  const size_t hw = (size_t)h * w;
  for (int i = threadIdx.x; i < kFN; i += kPhotoThreads) {
    const int ly = i / kFW, lx = i - ly * kFW;
    const size_t off = (size_t)pad_coord(ty0 + ly - 1, h) * w + pad_coord(tx0 + lx - 1, w);
#pragma unroll
    for (int c = 0; c < 3; ++c) tg[c * kFN + i] = ld1(target + ((size_t)bb * 3 + c) * hw + off);
  }
  for (int f = 0; f < 2; ++f) {
    __syncthreads();
    for (int i = threadIdx.x; i < kFN; i += kPhotoThreads) {
      const int ly = i / kFW, lx = i - ly * kFW;
      const size_t off = (size_t)pad_coord(ty0 + ly - 1, h) * w + pad_coord(tx0 + lx - 1, w);
#pragma unroll
      for (int c = 0; c < 3; ++c) sr[c * kFN + i] = ld1(sources + (((size_t)f * b + bb) * 3 + c) * hw + off);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int lx = threadIdx.x % kPhotoTW, ly = (threadIdx.x / kPhotoTW) * 2 + j;
      const int x = tx0 + lx, y = ty0 + ly;
      if (x < w && y < h) identity[((size_t)bb * 2 + f) * hw + (size_t)y * w + x] = reproj_error<kFW>(sr, tg, kFN, ly + 1, lx + 1);
    }
  }
}

// ---- forward --------------------------------------------------------------------------------
// One CTA = one 32x16 tile of one sample, looping over the scales: the target tile and its SSIM
// window statistics are loaded / computed once and reused by every scale and both source frames.
__device__ __forceinline__ void photo_fwd_body(const clslam_photo_args& a, int tiles_x, int tiles) {
  __shared__ float tg[3 * kFN];
  __shared__ float wx[2 * 3 * kFN];
  __shared__ float red[8];
  __shared__ float geo[16 + 24];           // inv_K (4x4), proj (2 x 3x4)
  __shared__ int up_i0[kFH + kFW], up_i1[kFH + kFW];     // bilinear up-sampling taps of the halo rows / columns
  __shared__ float up_l1[kFH + kFW];
  const int bb = blockIdx.y;
  const int h = a.h, w = a.w;
  const int tile = blockIdx.x;
  const int tx0 = (tile % tiles_x) * kPhotoTW, ty0 = (tile / tiles_x) * kPhotoTH;
  const size_t hw = (size_t)h * w;
  const int tid = threadIdx.x;
  if (tid < 16) geo[tid] = a.inv_k[bb * 16 + tid];
  else if (tid < 40) {
    const int i = tid - 16, f = i / 12;
    geo[tid] = a.proj[((size_t)f * a.b + bb) * 12 + (i - f * 12)];
  }
  for (int i = tid; i < kFN; i += kPhotoThreads) {
    const int ly = i / kFW, lx = i - ly * kFW;
    const size_t off = (size_t)pad_coord(ty0 + ly - 1, h) * w + pad_coord(tx0 + lx - 1, w);
#pragma unroll
    for (int c = 0; c < 3; ++c) tg[c * kFN + i] = ld1(a.target + ((size_t)bb * 3 + c) * hw + off);
  }
  __syncthreads();
  // this thread's two output pixels: (lx, 2*jy) and (lx, 2*jy + 1); target window sums are scale independent
  const int lx = tid % kPhotoTW, jy = tid / kPhotoTW;
  float sy[2][3], syy[2][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float rs[4], rss[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float* p = tg + c * kFN + (2 * jy + r) * kFW + lx;
      rs[r] = p[0] + p[1] + p[2];
      rss[r] = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    }
    sy[0][c] = rs[0] + rs[1] + rs[2]; syy[0][c] = rss[0] + rss[1] + rss[2];
    sy[1][c] = rs[1] + rs[2] + rs[3]; syy[1][c] = rss[1] + rss[2] + rss[3];
  }
  float local[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < a.n_scales; ++s) {
    const int hs = h >> s, ws = w >> s;
    const float* disp_b = a.disp[s] + (size_t)bb * hs * ws;
    float* depth_out = a.depth[s];
    float* warped_out = a.warped[s];
    if (tid < kFH + kFW) {
      const bool is_row = tid < kFH;
      const int coord = is_row ? pad_coord(ty0 + tid - 1, h) : pad_coord(tx0 + (tid - kFH) - 1, w);
      const UpCoef cf = is_row ? up_coef(coord, hs, h) : up_coef(coord, ws, w);
      up_i0[tid] = cf.i0; up_i1[tid] = cf.i1; up_l1[tid] = cf.l1;
    }
    __syncthreads();          // also: every read of wx by the previous scale is done
    for (int i = tid; i < kFN; i += kPhotoThreads) {
      const int ly = i / kFW, lxx = i - ly * kFW;
      const int py = ty0 + ly - 1, px = tx0 + lxx - 1;
      const int y = pad_coord(py, h), x = pad_coord(px, w);
      const size_t off = (size_t)y * w + x;
      const float ly1 = up_l1[ly], lx1 = up_l1[kFH + lxx];
      const float* r0 = disp_b + (size_t)up_i0[ly] * ws;
      const float* r1 = disp_b + (size_t)up_i1[ly] * ws;
      const int c0 = up_i0[kFH + lxx], c1 = up_i1[kFH + lxx];
      const float d = (1.f - ly1) * ((1.f - lx1) * ld1(r0 + c0) + lx1 * ld1(r0 + c1)) +
                      ly1 * ((1.f - lx1) * ld1(r1 + c0) + lx1 * ld1(r1 + c1));
      const float depth = depth_from_disp(d, a.depth_mode, a.min_depth, a.max_depth);
      const bool interior = (py == y) && (px == x) && ly >= 1 && ly <= kPhotoTH && lxx >= 1 && lxx <= kPhotoTW;
      if (interior && depth_out) depth_out[(size_t)bb * hw + off] = depth;
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        WarpGeom g;
        warp_geometry(geo, geo + 16 + f * 12, depth, x, y, h, w, g);
        const Taps t = make_taps(g.ix, g.iy, h, w);
        float v3[3];
        sample_rgba(a.sources_rgba + ((size_t)f * a.b + bb) * hw * 4, w, t, v3);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          wx[(f * 3 + c) * kFN + i] = v3[c];
          if (interior && warped_out) warped_out[(((size_t)f * a.b + bb) * 3 + c) * hw + off] = v3[c];
        }
      }
    }
    __syncthreads();
    // SSIM + L1 of both warped frames for the two pixels of this thread (shared row sums)
    float rep[2][2];          // [frame][pixel]
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      float ssim_sum[2] = {0.f, 0.f}, l1_sum[2] = {0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* xp = wx + (f * 3 + c) * kFN + (2 * jy) * kFW + lx;
        const float* yp = tg + c * kFN + (2 * jy) * kFW + lx;
        float rx[4], rxx[4], rxy[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float x0 = xp[r * kFW], x1 = xp[r * kFW + 1], x2 = xp[r * kFW + 2];
          const float y0 = yp[r * kFW], y1 = yp[r * kFW + 1], y2 = yp[r * kFW + 2];
          rx[r] = x0 + x1 + x2;
          rxx[r] = x0 * x0 + x1 * x1 + x2 * x2;
          rxy[r] = x0 * y0 + x1 * y1 + x2 * y2;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const SsimStats st = ssim_stats(rx[j] + rx[j + 1] + rx[j + 2], sy[j][c], rxx[j] + rxx[j + 1] + rxx[j + 2], syy[j][c],
                                          rxy[j] + rxy[j + 1] + rxy[j + 2]);
          ssim_sum[j] += clamp01(st.val);
          l1_sum[j] += fabsf(yp[(j + 1) * kFW + 1] - xp[(j + 1) * kFW + 1]);
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) rep[f][j] = 0.85f * (ssim_sum[j] * (1.f / 3.f)) + 0.15f * (l1_sum[j] * (1.f / 3.f));
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int x = tx0 + lx, y = ty0 + 2 * jy + j;
      if (x < w && y < h) {
        const size_t off = (size_t)y * w + x;
        float best = ld1(a.identity + ((size_t)bb * 2 + 0) * hw + off);
        float c1v = ld1(a.identity + ((size_t)bb * 2 + 1) * hw + off);
        if (a.noise[s]) {
          best += ld1(a.noise[s] + ((size_t)bb * 2 + 0) * hw + off) * 0.00001f;
          c1v += ld1(a.noise[s] + ((size_t)bb * 2 + 1) * hw + off) * 0.00001f;
        }
        int arg = 0;
        if (c1v < best) { best = c1v; arg = 1; }
        if (rep[0][j] < best) { best = rep[0][j]; arg = 2; }
        if (rep[1][j] < best) { best = rep[1][j]; arg = 3; }
        // mask_dynamic (:1064-1068): a potentially dynamic pixel is not selected -- no loss, and code 4 matches no gradient path
        if (a.dyn_mask && ld1(a.dyn_mask + (size_t)bb * hw + off) >= 0.5f) { best = 0.f; arg = 4; }
        a.argmin[s][(size_t)bb * hw + off] = (unsigned char)arg;
        local[s] += best;
      }
    }
  }
  for (int s = 0; s < a.n_scales; ++s) {
    const float total = block_sum(local[s], red);
    if (tid == 0) a.loss_partials[((size_t)s * a.b + bb) * tiles + tile] = total;
  }
}

__global__ void __launch_bounds__(kPhotoThreads)
photo_fwd_kernel(clslam_photo_args a, int tiles_x, int tiles) {
  pdl_wait();
  pdl_trigger();
  photo_fwd_body(a, tiles_x, tiles);
}
// EXPERIMENTAL variant (CLSLAM_PHOTO_OCC3=1): the same body capped at 85 registers, three CTAs per SM.  The warp phase is
// bound by the latency of its dependent gathers (ncu: long-scoreboard stalls, 25 % of the warp slots occupied)
__global__ void __launch_bounds__(kPhotoThreads, 3)
photo_fwd_occ3_kernel(clslam_photo_args a, int tiles_x, int tiles) {
  pdl_wait();
  pdl_trigger();
  photo_fwd_body(a, tiles_x, tiles);
}

// ---- backward -------------------------------------------------------------------------------
// The kernel body lives in photo_bwd_body.inc and is compiled once per variant.  The default kernel is the first one;
// the others are EXPERIMENTAL (CLSLAM_PHOTO_BWD_REUSE=1: phase 1 reads the forward pass's warped images;
// CLSLAM_PHOTO_OCC3=1: 85-register cap, three CTAs per SM)
#define PHOTO_BWD_NAME photo_bwd_kernel
#define PHOTO_BWD_BOUNDS __launch_bounds__(kPhotoThreads)
#define PHOTO_BWD_REUSE 0
#include "photo_bwd_body.inc"
#undef PHOTO_BWD_NAME
#undef PHOTO_BWD_BOUNDS
#undef PHOTO_BWD_REUSE

#define PHOTO_BWD_NAME photo_bwd_reuse_kernel
#define PHOTO_BWD_BOUNDS __launch_bounds__(kPhotoThreads)
#define PHOTO_BWD_REUSE 1
#include "photo_bwd_body.inc"
#undef PHOTO_BWD_NAME
#undef PHOTO_BWD_BOUNDS
#undef PHOTO_BWD_REUSE

#define PHOTO_BWD_NAME photo_bwd_occ3_kernel
#define PHOTO_BWD_BOUNDS __launch_bounds__(kPhotoThreads, 3)
#define PHOTO_BWD_REUSE 0
#include "photo_bwd_body.inc"
#undef PHOTO_BWD_NAME
#undef PHOTO_BWD_BOUNDS
#undef PHOTO_BWD_REUSE

#define PHOTO_BWD_NAME photo_bwd_reuse_occ3_kernel
#define PHOTO_BWD_BOUNDS __launch_bounds__(kPhotoThreads, 3)
#define PHOTO_BWD_REUSE 1
#include "photo_bwd_body.inc"
#undef PHOTO_BWD_NAME
#undef PHOTO_BWD_BOUNDS
#undef PHOTO_BWD_REUSE

// ---- transpose of the bilinear up-sampling + smoothness terms + sigmoid' -------------------
// scale 0 is the identity; for s >= 1 (f = 2^s) one thread owns one low-res pixel and walks the 2f x 2f full-res
// window whose bilinear taps (align_corners = False) touch it: rows f*r - f/2 .. f*r + 3f/2 - 1 (clipped)
__global__ void __launch_bounds__(256)
disp_grad_kernel(const float* __restrict__ ddisp_full, const float* __restrict__ disp, const float* __restrict__ smooth_grad,
                 int b_global, int has_sample0, int b, int h, int w, int scale, float* __restrict__ dz) {
  pdl_wait();
  pdl_trigger();
  const int hs = h >> scale, ws = w >> scale;
  const int f = 1 << scale;
  const size_t total = (size_t)b * hs * ws;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % ws);
    const int r = (int)((i / ws) % hs);
    const int bb = (int)(i / ((size_t)ws * hs));
    const float* src = ddisp_full + (size_t)bb * h * w;
    float g = 0.f;
    if (scale == 0) {
      g = src[(size_t)r * w + j];
    } else {
      const int y_lo = max(0, f * r - f / 2), y_hi = min(h, f * r + 3 * f / 2);
      const int x_lo = max(0, f * j - f / 2), x_hi = min(w, f * j + 3 * f / 2);
      for (int y = y_lo; y < y_hi; ++y) {
        const UpCoef cy = up_coef(y, hs, h);
        const float wy = (cy.i0 == r ? cy.l0 : 0.f) + (cy.i1 == r ? cy.l1 : 0.f);
        if (wy == 0.f) continue;
        float row = 0.f;
        for (int x = x_lo; x < x_hi; ++x) {
          const UpCoef cx = up_coef(x, ws, w);
          const float wxv = (cx.i0 == j ? cx.l0 : 0.f) + (cx.i1 == j ? cx.l1 : 0.f);
          row += wxv * ld1(src + (size_t)y * w + x);
        }
        g += wy * row;
      }
    }
    if (has_sample0 && bb == 0 && smooth_grad) {
      const float* sg = smooth_grad + (size_t)scale * (2 * (b_global + 1) + 1);
      g += sg[2 * (b_global + 1)];
      if (r < 2 && j <= b_global) g += sg[r * (b_global + 1) + j];
    }
    const float dv = disp[i];
    dz[i] = g * dv * (1.f - dv);
  }
}

// the same for the coarse scales (s >= 2), one WARP per low-res pixel: its 2f x 2f window (64 / 256 taps) is spread over the
// lanes and reduced with shuffles.  The thread-per-pixel form above leaves 1920 threads with 256 dependent taps each at scale 3
// of a 192 x 640 image (33 us at batch 1, on the lane that heads the weight-gradient chain).
__global__ void __launch_bounds__(256)
disp_grad_warp_kernel(const float* __restrict__ ddisp_full, const float* __restrict__ disp, const float* __restrict__ smooth_grad,
                      int b_global, int has_sample0, int b, int h, int w, int scale, float* __restrict__ dz) {
  pdl_wait();
  pdl_trigger();
  const int hs = h >> scale, ws = w >> scale;
  const int f = 1 << scale;
  const int lane = threadIdx.x & 31;
  const size_t total = (size_t)b * hs * ws;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < total; i += n_warps) {
    const int j = (int)(i % ws);
    const int r = (int)((i / ws) % hs);
    const int bb = (int)(i / ((size_t)ws * hs));
    const float* src = ddisp_full + (size_t)bb * h * w;
    const int y_lo = max(0, f * r - f / 2), y_hi = min(h, f * r + 3 * f / 2);
    const int x_lo = max(0, f * j - f / 2), x_hi = min(w, f * j + 3 * f / 2);
    const int nx = x_hi - x_lo, n = nx * (y_hi - y_lo);
    float g = 0.f;
    for (int k = lane; k < n; k += 32) {
      const int y = y_lo + k / nx, x = x_lo + k % nx;
      const UpCoef cy = up_coef(y, hs, h);
      const UpCoef cx = up_coef(x, ws, w);
      const float wy = (cy.i0 == r ? cy.l0 : 0.f) + (cy.i1 == r ? cy.l1 : 0.f);
      const float wxv = (cx.i0 == j ? cx.l0 : 0.f) + (cx.i1 == j ? cx.l1 : 0.f);
      g += wy * wxv * ld1(src + (size_t)y * w + x);
    }
    g = warp_sum(g);
    if (lane == 0) {
      if (has_sample0 && bb == 0 && smooth_grad) {
        const float* sg = smooth_grad + (size_t)scale * (2 * (b_global + 1) + 1);
        g += sg[2 * (b_global + 1)];
        if (r < 2 && j <= b_global) g += sg[r * (b_global + 1) + j];
      }
      const float dv = disp[i];
      dz[i] = g * dv * (1.f - dv);
    }
  }
}

__global__ void __launch_bounds__(256)
disp_to_depth_kernel(const float* __restrict__ disp, float* __restrict__ depth, long long n, int mode, float min_depth, float max_depth) {
  pdl_wait();
  pdl_trigger();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    depth[i] = depth_from_disp(ld1(disp + i), mode, min_depth, max_depth);
}

}  // namespace clslam

using namespace clslam;

extern "C" {

int clslam_pack_rgba(const float* planar, float* packed, int n_images, int h, int w, void* stream) {
  const size_t total = (size_t)n_images * h * w;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(pack_rgba_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), planar, packed, n_images, h * w);
  return finish_launch("pack_rgba");
}

int clslam_photo_tiles(int h, int w) { return ((w + kPhotoTW - 1) / kPhotoTW) * ((h + kPhotoTH - 1) / kPhotoTH); }

int clslam_photo_identity(const float* target, const float* sources, float* identity, int b, int h, int w, void* stream) {
  if (h < 2 || w < 2) { set_error("photo_identity: image too small"); return -1; }
  const int tiles_x = (w + kPhotoTW - 1) / kPhotoTW;
  dim3 grid(clslam_photo_tiles(h, w), b);
  launch_k(photo_identity_kernel, dim3(grid), dim3(kPhotoThreads), 0, as_stream(stream), target, sources, identity, b, h, w, tiles_x);
  return finish_launch("photo_identity");
}

static int check_photo(const clslam_photo_args* a, const char* what) {
  if (!a) { set_error("%s: null args", what); return -1; }
  if (!a->sources_rgba) { set_error("%s: sources_rgba is null (clslam_pack_rgba)", what); return -1; }
  if (a->n_scales < 1 || a->n_scales > 4 || a->b < 1 || a->h < 2 || a->w < 2) { set_error("%s: bad dims", what); return -1; }
  if ((a->h % (1 << (a->n_scales - 1))) || (a->w % (1 << (a->n_scales - 1)))) { set_error("%s: h, w must be divisible by 2^(n_scales-1)", what); return -1; }
  if ((long long)a->b * a->n_scales > 65535) { set_error("%s: b * n_scales > 65535", what); return -1; }
  return 0;
}

// kernel variants (A/B-tested on hardware in round 2, both on by default; NAME=0 in the environment selects the round-1 kernels)
static bool env_flag(const char* name) {
  const char* e = getenv(name);
  return !e || e[0] != '0';
}
static bool photo_occ3() { static const bool v = env_flag("CLSLAM_PHOTO_OCC3"); return v; }
static bool photo_bwd_reuse() { static const bool v = env_flag("CLSLAM_PHOTO_BWD_REUSE"); return v; }

int clslam_photo_fwd(const clslam_photo_args* a, void* stream) {
  if (int rc = check_photo(a, "photo_fwd")) return rc;
  const int tiles_x = (a->w + kPhotoTW - 1) / kPhotoTW, tiles = clslam_photo_tiles(a->h, a->w);
  dim3 grid(tiles, a->b);
  launch_k(photo_occ3() ? photo_fwd_occ3_kernel : photo_fwd_kernel, dim3(grid), dim3(kPhotoThreads), 0, as_stream(stream), *a,
           tiles_x, tiles);
  return finish_launch("photo_fwd");
}

int clslam_photo_bwd(const clslam_photo_args* a, void* stream) {
  if (int rc = check_photo(a, "photo_bwd")) return rc;
  const int tiles_x = (a->w + kPhotoTW - 1) / kPhotoTW, tiles = clslam_photo_tiles(a->h, a->w);
  dim3 grid(tiles, a->b * a->n_scales);
  bool reuse = photo_bwd_reuse();          // needs the forward pass's warped images of every scale
  for (int s = 0; s < a->n_scales; ++s) reuse = reuse && a->warped[s] != nullptr;
  auto kern = reuse ? (photo_occ3() ? photo_bwd_reuse_occ3_kernel : photo_bwd_reuse_kernel)
                    : (photo_occ3() ? photo_bwd_occ3_kernel : photo_bwd_kernel);
  launch_k(kern, dim3(grid), dim3(kPhotoThreads), 0, as_stream(stream), *a, tiles_x, tiles);
  return finish_launch("photo_bwd");
}

int clslam_disp_to_depth(const float* disp, float* depth, long long n, int depth_mode, float min_depth, float max_depth, void* stream) {
  if (!disp || !depth || n < 1) { set_error("disp_to_depth: bad arguments"); return -1; }
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(disp_to_depth_kernel, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream), disp, depth, n, depth_mode, min_depth, max_depth);
  return finish_launch("disp_to_depth");
}

int clslam_disp_grad(const float* ddisp_full, const float* disp, const float* smooth_grad, int b_global, int has_sample0,
                     int b, int h, int w, int scale, float* dz_disp, void* stream) {
  const size_t total = (size_t)b * (h >> scale) * (w >> scale);
  if (scale >= 2) {          // one warp per low-res pixel
    size_t blocks = (total + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_k(disp_grad_warp_kernel, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream), ddisp_full, disp, smooth_grad, b_global,
             has_sample0, b, h, w, scale, dz_disp);
    return finish_launch("disp_grad");
  }
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(disp_grad_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), ddisp_full, disp, smooth_grad, b_global, has_sample0, b, h, w, scale,
                                                         dz_disp);
  return finish_launch("disp_grad");
}

}  // extern "C"
