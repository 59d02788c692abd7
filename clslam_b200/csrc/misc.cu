// Memory-bound helpers of the conv stack: image staging, max-pool, the one-channel dispconv
// (forward / wgrad / dgrad) and the flat Adam update.
#include "conv_ops.cuh"
#include "planes.cuh"

namespace clslam {

// ---- (x - 0.45) / 0.225, NCHW -> NHWC, channel padding ------------------------------------
__global__ void __launch_bounds__(256)
stage_images_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n, int hw,
                    int c_pad) {
  pdl_wait();
  pdl_trigger();
  const size_t total = (size_t)n * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t img = i / hw, px = i - img * hw;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = 0.f;
    const float* pa = a + img * 3 * hw + px;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (ld1(pa + (size_t)c * hw) - 0.45f) / 0.225f;
    if (b) {
      const float* pb = b + img * 3 * hw + px;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[3 + c] = (ld1(pb + (size_t)c * hw) - 0.45f) / 0.225f;
    }
    float4* o = reinterpret_cast<float4*>(out + i * c_pad);
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    if (c_pad == 8) o[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// ---- 3x3 / s2 / p1 max pool, NHWC ----------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int hi, int wi, int c, int ho, int wo, PlaneView po) {
  pdl_wait();
  pdl_trigger();
  const int c4 = c / 4;
  const size_t total = (size_t)n * ho * wo * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c4) * 4;
    size_t p = i / c4;
    const int ow = (int)(p % wo); p /= wo;
    const int oh = (int)(p % ho);
    const int b = (int)(p / ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int r = 0; r < 3; ++r) {
      const int ih = oh * 2 + r - 1;
      if (ih < 0 || ih >= hi) continue;
      for (int s = 0; s < 3; ++s) {
        const int iw = ow * 2 + s - 1;
        if (iw < 0 || iw >= wi) continue;
        const float4 v = ld4(x + (((size_t)b * hi + ih) * wi + iw) * c + cc);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    if (y) *reinterpret_cast<float4*>(y + i * 4) = m;
    if (po.base) {
      const float v[4] = {m.x, m.y, m.z, m.w};
      planes_emit<4>(po, b, oh, ow, cc, v);
    }
  }
}

// work index -> pixel, walking 8 x 16 pixel tiles instead of image rows: the 3x3 taps of the pixels a block touches in
// one sweep then share their rows through L1 instead of re-reading them from L2 (falls back to row order for odd sizes)
__device__ __forceinline__ void tile_order(size_t p, int h, int wd, int& b, int& hh, int& ww) {
  const int hw = h * wd;
  b = (int)(p / (size_t)hw);
  const int r = (int)(p - (size_t)b * hw);
  if ((h & 7) == 0 && (wd & 15) == 0) {
    const int tile = r >> 7, off = r & 127, tiles_x = wd >> 4;
    const int ty = tile / tiles_x;
    hh = ty * 8 + (off >> 4);
    ww = (tile - ty * tiles_x) * 16 + (off & 15);
  } else {
    hh = r / wd;
    ww = r - hh * wd;
  }
}

// ---- dispconv forward: reflect 3x3 -> 1 channel, + bias, sigmoid ---------------------------
// G = ci/4 lanes cooperate on one pixel (each owns 4 channels), 32/G pixels per warp.
template <int G>
__global__ void __launch_bounds__(256)
dispconv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ y, int n, int h, int wd) {
  pdl_wait();
  pdl_trigger();
  constexpr int CI = G * 4;
  constexpr int PPW = 32 / G;
  const int lane = threadIdx.x & 31, sub = lane % G, grp = lane / G;
  float4 wr[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wr[t] = ld4(w + t * CI + sub * 4);
  const float bv = ld1(bias);
  const size_t total = (size_t)n * h * wd;
  const size_t warp_id = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t base = warp_id * PPW; base < total; base += n_warps * PPW) {
    const size_t p = base + grp;
    float acc = 0.f;
    int b = 0, hh = 0, ww = 0;
    if (p < total) {
      tile_order(p, h, wd, b, hh, ww);
      // one 64-bit base address per pixel, 32-bit row / column offsets per tap (the round-1 kernel redid the full 64-bit index
      // arithmetic for each of the nine taps: ~1400 thread-instructions per pixel); all nine loads are issued
      // before the first multiply-add
      const float* xb = x + (((size_t)b * h + hh) * wd + ww) * CI + sub * 4;
      int dr[3], dc[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        dr[r] = (reflect_idx(hh + r - 1, h) - hh) * wd * CI;
        dc[r] = (reflect_idx(ww + r - 1, wd) - ww) * CI;
      }
      float4 v[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) v[r * 3 + s] = ld4(xb + dr[r] + dc[s]);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 k = wr[t];
        acc = fmaf(v[t].x, k.x, acc); acc = fmaf(v[t].y, k.y, acc); acc = fmaf(v[t].z, k.z, acc); acc = fmaf(v[t].w, k.w, acc);
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (sub == 0 && p < total) y[((size_t)b * h + hh) * wd + ww] = act_apply(acc + bv, CLSLAM_ACT_SIGMOID);
  }
}

// ---- dispconv wgrad: dW[tap][ci] = sum_p dz[p] * x_reflect[p + tap][ci], db = sum_p dz[p] ---
// Evaluated from the INPUT side: dW[tap][c] = sum_q x[q][c] * G[q][tap] with G[q][tap] = sum of dz over the output pixels
// that read q through tap (one pixel in the interior; the reflection border adds the mirrored ones, dgrad_sources_1d), so
// every x[q] is loaded exactly once (row-major, 512 contiguous bytes per warp) and the nine dz values come from L1
// (dz has one channel).  The round-1 kernel gathered x nine times per pixel and ran at 3.5 % of the HBM roofline.
// per-block partials -> ws[block][9*ci + 1]; reduced in fixed order by dispconv_wgrad_reduce (deterministic).
template <int G>
__global__ void __launch_bounds__(256)
dispconv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz, float* __restrict__ ws, int n, int h,
                      int wd) {
  pdl_wait();
  pdl_trigger();
  constexpr int CI = G * 4;
  constexpr int PPW = 32 / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % G, grp = lane / G;
  float4 acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = f4zero();
  float bsum = 0.f;
  const int hw = h * wd;
  const size_t total = (size_t)n * hw;
  const size_t warp_id = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t base = warp_id * PPW; base < total; base += n_warps * PPW) {
    const size_t p = base + grp;
    if (p >= total) continue;
    const int b = (int)(p / (size_t)hw);
    const int rem = (int)(p - (size_t)b * hw);
    const int hh = rem / wd, ww = rem - hh * wd;
    const float4 v = ld4(x + p * CI + sub * 4);
    const float* dzb = dz + (size_t)b * hw;
    float g[9];
    if (hh >= 2 && hh < h - 2 && ww >= 2 && ww < wd - 2) {          // interior: the output pixel q - (tap - 1)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) g[r * 3 + s2] = ld1(dzb + (hh + 1 - r) * wd + (ww + 1 - s2));
    } else {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) {
          float gs = 0.f;
          dgrad_sources_1d(hh, r, 1, h, 1, [&](int sr) {
            dgrad_sources_1d(ww, s2, 1, wd, 1, [&](int sc) { gs += ld1(dzb + sr * wd + sc); });
          });
          g[r * 3 + s2] = gs;
        }
    }
    if (sub == 0) bsum += ld1(dzb + rem);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float4& a = acc[t];
      a.x = fmaf(g[t], v.x, a.x); a.y = fmaf(g[t], v.y, a.y); a.z = fmaf(g[t], v.z, a.z); a.w = fmaf(g[t], v.w, a.w);
    }
  }
  // reduce the PPW pixel groups of each warp (lanes with equal `sub`)
#pragma unroll
  for (int t = 0; t < 9; ++t) {
#pragma unroll
    for (int o = G; o < 32; o <<= 1) {
      acc[t].x += __shfl_xor_sync(0xffffffffu, acc[t].x, o);
      acc[t].y += __shfl_xor_sync(0xffffffffu, acc[t].y, o);
      acc[t].z += __shfl_xor_sync(0xffffffffu, acc[t].z, o);
      acc[t].w += __shfl_xor_sync(0xffffffffu, acc[t].w, o);
    }
  }
  bsum = warp_sum(bsum);
  __shared__ float red[8][9 * CI + 1];
  if (lane < G) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float* d = &red[warp][t * CI + lane * 4];
      d[0] = acc[t].x; d[1] = acc[t].y; d[2] = acc[t].z; d[3] = acc[t].w;
    }
  }
  if (lane == 0) red[warp][9 * CI] = bsum;
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * CI + 1; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv][i];
    ws[(size_t)blockIdx.x * (9 * CI + 1) + i] = s;
  }
}

// one block = 32 outputs x 8 warps: warp g adds the partial sums of blocks g, g + 8, ... (coalesced over the 32 outputs), the
// eight chains are combined in a fixed order (deterministic).  The round-1 kernel gave every output ONE thread that walked
// all blocks serially: 50 us for 9 * ci + 1 <= 1153 outputs.
__global__ void __launch_bounds__(256)
dispconv_wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, float* __restrict__ db, int blocks,
                             int ci) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8][33];
  const int len = 9 * ci + 1;
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float a0 = 0.f, a1 = 0.f;
  if (i < len) {
    int b = g;
    for (; b + 8 < blocks; b += 16) { a0 += ws[(size_t)b * len + i]; a1 += ws[(size_t)(b + 8) * len + i]; }
    if (b < blocks) a0 += ws[(size_t)b * len + i];
  }
  red[g][lane] = a0 + a1;
  __syncthreads();
  if (g == 0 && i < len) {
    const float s = ((red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane])) +
                    ((red[4][lane] + red[5][lane]) + (red[6][lane] + red[7][lane]));
    if (i < 9 * ci) dw[i] = s; else if (db) db[0] = s;
  }
}

// ---- dispconv dgrad when it is the only consumer: dz_prev = act'(y_prev) * dX ---------------
__global__ void __launch_bounds__(256)
dispconv_dgrad_act_kernel(const float* __restrict__ dz, const float* __restrict__ w, const float* __restrict__ y_prev,
                          int act_prev, float* __restrict__ dz_prev, int n, int h, int wd, int ci, PlaneView po) {
  pdl_wait();
  pdl_trigger();
  const int c4 = ci / 4;
  const size_t total = (size_t)n * h * wd * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c4) * 4;
    size_t p = i / c4;
    const int ww = (int)(p % wd); p /= wd;
    const int hh = (int)(p % h);
    const int b = (int)(p / h);
    float4 g = dispconv_dgrad_at(dz, w, b, hh, ww, h, wd, ci, cc);
    const float4 yv = ld4(y_prev + i * 4);
    g.x *= act_grad_from_output(yv.x, act_prev);
    g.y *= act_grad_from_output(yv.y, act_prev);
    g.z *= act_grad_from_output(yv.z, act_prev);
    g.w *= act_grad_from_output(yv.w, act_prev);
    *reinterpret_cast<float4*>(dz_prev + i * 4) = g;
    if (po.base) {
      const float v[4] = {g.x, g.y, g.z, g.w};
      planes_emit<4>(po, b, hh, ww, cc, v);
    }
  }
}

// ---- Adam (torch.optim.Adam single-tensor math, no weight decay / amsgrad) ------------------
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, double lr, double beta1, double beta2, double eps_d, int step_host,
            const int* __restrict__ step_dev) {
  pdl_wait();
  pdl_trigger();
  // step_dev holds the number of updates done BEFORE this one (adam_tick_kernel bumps it after).
  // The bias corrections need double-precision pow: one thread per block computes them, the rest read shared memory.
  __shared__ float s_neg_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const int step = step_dev ? (*step_dev + 1) : step_host;
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    s_neg_step_size = (float)(-(lr / bc1));
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  __syncthreads();
  const float neg_step_size = s_neg_step_size, bc2_sqrt = s_bc2_sqrt;
  const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2), eps = (float)eps_d;
  const long long n4 = n >> 2;                           // the flat buffers are padded to multiples of 64 elements
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = ld4(g + i * 4);
    float4 m4 = *reinterpret_cast<const float4*>(m + i * 4), v4 = *reinterpret_cast<const float4*>(v + i * 4);
    float4 p4 = *reinterpret_cast<const float4*>(p + i * 4);
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mm[j] = mm[j] + w1 * (gg[j] - mm[j]);              // exp_avg.lerp_(grad, 1 - beta1)
      vv[j] = vv[j] * b2 + w2 * gg[j] * gg[j];           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(vv[j]) / bc2_sqrt + eps;
      pp[j] = pp[j] + neg_step_size * mm[j] / denom;     // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
    *reinterpret_cast<float4*>(p + i * 4) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *reinterpret_cast<float4*>(m + i * 4) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i * 4) = make_float4(vv[0], vv[1], vv[2], vv[3]);
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float mi = m[i], vi = v[i];
    mi = mi + w1 * (gi - mi);
    vi = vi * b2 + w2 * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] + neg_step_size * mi / denom;
    m[i] = mi; v[i] = vi;
  }
}
__global__ void adam_tick_kernel(int* step_dev) {
  pdl_wait();
  pdl_trigger(); *step_dev += 1; }

static int grid_for(size_t work, int per_block = 256, int cap = 148 * 8) {
  size_t b = (work + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > (size_t)cap) b = cap;
  return (int)b;
}

}  // namespace clslam

using namespace clslam;

extern "C" {

int clslam_stage_images(const float* img_a, const float* img_b, float* out_nhwc, int n, int h, int w, int c_pad,
                        void* stream) {
  if (!(c_pad == 4 || c_pad == 8) || (img_b && c_pad != 8)) { set_error("stage_images: c_pad must be 4 (one image) or 8 (two)"); return -1; }
  launch_k(stage_images_kernel, dim3(grid_for((size_t)n * h * w)), dim3(256), 0, as_stream(stream), img_a, img_b, out_nhwc, n, h * w, c_pad);
  return finish_launch("stage_images");
}

int clslam_maxpool3x3s2_emit(const float* x, float* y, int n, int hi, int wi, int c, const clslam_planes* out, int out_halo,
                             void* stream) {
  if (c % 4) { set_error("maxpool: c %% 4 != 0"); return -1; }
  const int ho = (hi + 2 - 3) / 2 + 1, wo = (wi + 2 - 3) / 2 + 1;
  PlaneView po{};
  if (out) {
    if (!out->base || out->n != n || out->h != ho || out->w != wo || out->c != c || out->c_total % 8 || out->c_off % 4) {
      set_error("maxpool: plane output does not match (%d x %d x %d)", ho, wo, c); return -1;
    }
    po = make_plane_view(out, 1, out_halo);
  }
  if (!y && !po.base) { set_error("maxpool: no output"); return -1; }
  launch_k(maxpool_kernel, dim3(grid_for((size_t)n * ho * wo * (c / 4))), dim3(256), 0, as_stream(stream), x, y, n, hi, wi, c, ho, wo, po);
  return finish_launch("maxpool");
}

int clslam_maxpool3x3s2(const float* x, float* y, int n, int hi, int wi, int c, void* stream) {
  return clslam_maxpool3x3s2_emit(x, y, n, hi, wi, c, nullptr, 0, stream);
}

int clslam_dispconv_fwd(const float* x, const float* w, const float* bias, float* y, int n, int h, int w_, int ci,
                        void* stream) {
  const size_t px = (size_t)n * h * w_;
  cudaStream_t st = as_stream(stream);
  switch (ci) {
    case 16: launch_k(dispconv_fwd_kernel<4>, dim3(grid_for(px, 64)), dim3(256), 0, st, x, w, bias, y, n, h, w_); break;
    case 32: launch_k(dispconv_fwd_kernel<8>, dim3(grid_for(px, 32)), dim3(256), 0, st, x, w, bias, y, n, h, w_); break;
    case 64: launch_k(dispconv_fwd_kernel<16>, dim3(grid_for(px, 16)), dim3(256), 0, st, x, w, bias, y, n, h, w_); break;
    case 128: launch_k(dispconv_fwd_kernel<32>, dim3(grid_for(px, 8)), dim3(256), 0, st, x, w, bias, y, n, h, w_); break;
    default: set_error("dispconv_fwd: ci must be 16/32/64/128 (got %d)", ci); return -1;
  }
  return finish_launch("dispconv_fwd");
}

static int dispconv_wgrad_blocks(int n, int h, int w_, int ci) {
  const size_t px = (size_t)n * h * w_;
  const int ppb = 8 * (32 / (ci / 4));          // pixels per block-iteration
  size_t b = (px + (size_t)ppb * 8 - 1) / ((size_t)ppb * 8);
  if (b < 1) b = 1;
  if (b > 444) b = 444;          // 3 resident blocks per SM (78 registers): one wave; the partial sums are reduced in block order
  return (int)b;
}

long long clslam_dispconv_wgrad_workspace_bytes(int n, int h, int w_, int ci) {
  return (long long)dispconv_wgrad_blocks(n, h, w_, ci) * (9 * ci + 1) * 4;
}

int clslam_dispconv_wgrad(const float* x, const float* dz, float* dw, float* db, int n, int h, int w_, int ci,
                          void* workspace, long long workspace_bytes, void* stream) {
  const int blocks = dispconv_wgrad_blocks(n, h, w_, ci);
  if (!workspace || workspace_bytes < (long long)blocks * (9 * ci + 1) * 4) { set_error("dispconv_wgrad: workspace too small"); return -2; }
  float* ws = (float*)workspace;
  cudaStream_t st = as_stream(stream);
  switch (ci) {
    case 16: launch_k(dispconv_wgrad_kernel<4>, dim3(blocks), dim3(256), 0, st, x, dz, ws, n, h, w_); break;
    case 32: launch_k(dispconv_wgrad_kernel<8>, dim3(blocks), dim3(256), 0, st, x, dz, ws, n, h, w_); break;
    case 64: launch_k(dispconv_wgrad_kernel<16>, dim3(blocks), dim3(256), 0, st, x, dz, ws, n, h, w_); break;
    case 128: launch_k(dispconv_wgrad_kernel<32>, dim3(blocks), dim3(256), 0, st, x, dz, ws, n, h, w_); break;
    default: set_error("dispconv_wgrad: ci must be 16/32/64/128 (got %d)", ci); return -1;
  }
  if (int rc = finish_launch("dispconv_wgrad")) return rc;
  launch_k(dispconv_wgrad_reduce_kernel, dim3((9 * ci + 1 + 31) / 32), dim3(256), 0, st, ws, dw, db, blocks, ci);
  return finish_launch("dispconv_wgrad_reduce");
}

int clslam_dispconv_dgrad_act_emit(const float* dz, const float* w, const float* y_prev, int act_prev, float* dz_prev,
                                   const clslam_planes* dz_prev_planes, int n, int h, int w_, int ci, void* stream) {
  if (ci % 4) { set_error("dispconv_dgrad: ci %% 4 != 0"); return -1; }
  PlaneView po{};
  if (dz_prev_planes) {
    const clslam_planes* q = dz_prev_planes;
    if (!q->base || q->n != n || q->h != h || q->w != w_ || q->c != ci || q->c_total % 8 || q->c_off % 4) {
      set_error("dispconv_dgrad: planes mismatch"); return -1;
    }
    po = make_plane_view(q, 1, 0);
  }
  launch_k(dispconv_dgrad_act_kernel, dim3(grid_for((size_t)n * h * w_ * (ci / 4))), dim3(256), 0, as_stream(stream), dz, w, y_prev, act_prev,
                                                                                                 dz_prev, n, h, w_, ci, po);
  return finish_launch("dispconv_dgrad_act");
}

int clslam_dispconv_dgrad_act(const float* dz, const float* w, const float* y_prev, int act_prev, float* dz_prev, int n,
                              int h, int w_, int ci, void* stream) {
  return clslam_dispconv_dgrad_act_emit(dz, w, y_prev, act_prev, dz_prev, nullptr, n, h, w_, ci, stream);
}

int clslam_adam_tick(int* step_dev, void* stream) {
  if (!step_dev) { set_error("adam_tick: null counter"); return -1; }
  launch_k(adam_tick_kernel, dim3(1), dim3(1), 0, as_stream(stream), step_dev);
  return finish_launch("adam_tick");
}

int clslam_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, double lr,
                     double beta1, double beta2, double eps, int step, int* step_dev, void* stream) {
  cudaStream_t st = as_stream(stream);
  launch_k(adam_kernel, dim3(grid_for((size_t)n, 256 * 4)), dim3(256), 0, st, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step,
                                                           step_dev);
  if (int rc = finish_launch("adam")) return rc;
  if (step_dev && step != CLSLAM_ADAM_NO_TICK) {
    launch_k(adam_tick_kernel, dim3(1), dim3(1), 0, st, step_dev);
    return finish_launch("adam_tick");
  }
  return 0;
}

}  // extern "C"
