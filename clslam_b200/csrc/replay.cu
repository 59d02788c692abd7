// Replay-sample image path on the device: what slam/replay_buffer.py:263-291 does per replay
// sample on the CPU -- PIL LANCZOS resize to the network size, the progressive half-size pyramid, ToTensor, and the colour
// jitter of datasets/utils.py:236-259 (torchvision tensor ops) -- as memory-bound kernels over a decoded uint8 image.
//
//   * resampling: Pillow's ImagingResample for 8-bit images, bit-exact.  Coefficients are fixed-point integers
//     (PRECISION_BITS = 22) built on the host exactly as Pillow's precompute_coeffs / normalize_coeffs_8bpc do
//     (clslam_b200/replay_path.py); the kernels are the two integer passes (horizontal, then vertical).
//   * ToTensor: HWC uint8 -> CHW float / 255.
//   * colour jitter: torchvision.transforms.functional adjust_{brightness, contrast, saturation, hue} on float tensors,
//     applied in the sampled order; products and sums are rounded separately (no FMA contraction), as separate torch kernels do.
#include <stdint.h>

#include "common.cuh"

namespace clslam {

constexpr int kResPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8_fixed(int v) {
  v >>= kResPrecisionBits;                      // arithmetic shift, as Pillow's clip8 lookup index
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// dst[n][y][xo][c] = clip8(2^21 + sum_k src[n][y][xmin(xo) + k][c] * kk[xo][k])
__global__ void __launch_bounds__(256)
resample_u8_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n, int h, int w_in, int w_out, int c,
                     const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  pdl_wait();
  pdl_trigger();
  const size_t total = (size_t)n * h * w_out * c;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    size_t p = i / c;
    const int xo = (int)(p % w_out);
    const size_t row = p / w_out;               // n * h + y
    const int xmin = bounds[2 * xo], xmax = bounds[2 * xo + 1];
    const int* k = kk + (size_t)xo * ksize;
    const uint8_t* line = src + (row * w_in + xmin) * c + ch;
    int ss = 1 << (kResPrecisionBits - 1);
    for (int x = 0; x < xmax; ++x) ss += (int)line[(size_t)x * c] * k[x];
    dst[i] = clip8_fixed(ss);
  }
}

// dst[n][yo][x][c] = clip8(2^21 + sum_k src[n][ymin(yo) + k][x][c] * kk[yo][k])
__global__ void __launch_bounds__(256)
resample_u8_v_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n, int h_in, int h_out, int w, int c,
                     const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  pdl_wait();
  pdl_trigger();
  const size_t line = (size_t)w * c;
  const size_t total = (size_t)n * h_out * line;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t xc = i % line;
    size_t p = i / line;
    const int yo = (int)(p % h_out);
    const size_t img = p / h_out;
    const int ymin = bounds[2 * yo], ymax = bounds[2 * yo + 1];
    const int* k = kk + (size_t)yo * ksize;
    const uint8_t* col = src + (img * h_in + ymin) * line + xc;
    int ss = 1 << (kResPrecisionBits - 1);
    for (int y = 0; y < ymax; ++y) ss += (int)col[(size_t)y * line] * k[y];
    dst[i] = clip8_fixed(ss);
  }
}

// torchvision ToTensor: (n, h, w, 3) uint8 -> (n, 3, h, w) float32 / 255
__global__ void __launch_bounds__(256)
u8_to_f32_chw_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int n, int hw) {
  pdl_wait();
  pdl_trigger();
  const size_t total = (size_t)n * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t img = i / hw, px = i - img * hw;
    const uint8_t* s = src + i * 3;
    float* d = dst + img * 3 * hw + px;
    d[0] = __fdiv_rn((float)s[0], 255.f);
    d[(size_t)hw] = __fdiv_rn((float)s[1], 255.f);
    d[2 * (size_t)hw] = __fdiv_rn((float)s[2], 255.f);
  }
}

// ---- colour jitter (torchvision/transforms/_functional_tensor.py) ---------------------------------------------------
__device__ __forceinline__ float gray_of(float r, float g, float b) {      // rgb_to_grayscale: (0.2989 r + 0.587 g) + 0.114 b
  return __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
}
__device__ __forceinline__ float blend1(float a, float b, float ratio, float inv) {      // _blend, float images: bound 1.0
  const float v = __fadd_rn(__fmul_rn(ratio, a), __fmul_rn(inv, b));
  return fminf(fmaxf(v, 0.f), 1.f);
}

// sum of the grey values of every image (adjust_contrast's mean): per-block partial sums, fixed-order final sum
constexpr int kGrayBlocks = 64;
__global__ void __launch_bounds__(256)
gray_partial_kernel(const float* __restrict__ img, int hw, double* __restrict__ part) {
  pdl_wait();
  pdl_trigger();
  const int im = blockIdx.y;
  const float* r = img + (size_t)im * 3 * hw;
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x)
    acc += (double)gray_of(r[i], r[(size_t)hw + i], r[2 * (size_t)hw + i]);
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[(size_t)im * gridDim.x + blockIdx.x] = red[0];
}

// op: 0 brightness, 1 contrast, 2 saturation, 3 hue
__global__ void __launch_bounds__(256)
jitter_op_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int hw, int op, float ratio, float inv,
                 const double* __restrict__ gray_part, int gray_blocks) {
  pdl_wait();
  pdl_trigger();
  const size_t total = (size_t)n * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t im = i / hw, px = i - im * hw;
    const float* s = src + im * 3 * hw + px;
    float* d = dst + im * 3 * hw + px;
    float r = s[0], g = s[(size_t)hw], b = s[2 * (size_t)hw];
    if (op == 0) {
      r = blend1(r, 0.f, ratio, inv); g = blend1(g, 0.f, ratio, inv); b = blend1(b, 0.f, ratio, inv);
    } else if (op == 1) {
      double sum = 0.0;
      for (int k = 0; k < gray_blocks; ++k) sum += gray_part[im * gray_blocks + k];
      const float mean = (float)(sum / (double)hw);
      r = blend1(r, mean, ratio, inv); g = blend1(g, mean, ratio, inv); b = blend1(b, mean, ratio, inv);
    } else if (op == 2) {
      const float l = gray_of(r, g, b);
      r = blend1(r, l, ratio, inv); g = blend1(g, l, ratio, inv); b = blend1(b, l, ratio, inv);
    } else {
      // _rgb2hsv
      const float maxc = fmaxf(r, fmaxf(g, b)), minc = fminf(r, fminf(g, b));
      const bool eqc = maxc == minc;
      const float cr = __fsub_rn(maxc, minc);
      const float sat = __fdiv_rn(cr, eqc ? 1.f : maxc);
      const float div = eqc ? 1.f : cr;
      const float rc = __fdiv_rn(__fsub_rn(maxc, r), div), gc = __fdiv_rn(__fsub_rn(maxc, g), div), bc = __fdiv_rn(__fsub_rn(maxc, b), div);
      const float hr = (maxc == r) ? __fsub_rn(bc, gc) : 0.f;
      const float hg = (maxc == g && maxc != r) ? __fsub_rn(__fadd_rn(2.f, rc), bc) : 0.f;
      const float hb = (maxc != g && maxc != r) ? __fsub_rn(__fadd_rn(4.f, gc), rc) : 0.f;
      float hh = __fadd_rn(__fadd_rn(hr, hg), hb);
      hh = fmodf(__fadd_rn(__fdiv_rn(hh, 6.f), 1.f), 1.f);
      // h = (h + hue_factor) % 1.0 (floor modulus), `ratio` carries hue_factor
      hh = __fadd_rn(hh, ratio);
      hh = __fsub_rn(hh, floorf(hh));
      if (hh >= 1.f) hh = 0.f;                          // remainder(x, 1) lies in [0, 1)
      // _hsv2rgb
      const float h6 = __fmul_rn(hh, 6.f);
      const float fl = floorf(h6);
      const float f = __fsub_rn(h6, fl);
      int sector = (int)fl % 6;
      if (sector < 0) sector += 6;
      const float v = maxc;
      const float p = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, sat)), 0.f), 1.f);
      const float q = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, __fmul_rn(sat, f))), 0.f), 1.f);
      const float t = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, __fmul_rn(sat, __fsub_rn(1.f, f)))), 0.f), 1.f);
      switch (sector) {
        case 0: r = v; g = t; b = p; break;
        case 1: r = q; g = v; b = p; break;
        case 2: r = p; g = v; b = t; break;
        case 3: r = p; g = q; b = v; break;
        case 4: r = t; g = p; b = v; break;
        default: r = v; g = p; b = q; break;
      }
    }
    d[0] = r; d[(size_t)hw] = g; d[2 * (size_t)hw] = b;
  }
}

static int grid_1d(size_t work, int cap = 148 * 8) {
  size_t b = (work + 255) / 256;
  if (b < 1) b = 1;
  if (b > (size_t)cap) b = cap;
  return (int)b;
}

}  // namespace clslam

using namespace clslam;

extern "C" {

int clslam_resample_u8(const void* src, void* dst, int n, int h_in, int w_in, int c, int out_size, int vertical, const int* bounds,
                       const int* kk, int ksize, void* stream) {
  if (!src || !dst || !bounds || !kk || n < 1 || h_in < 1 || w_in < 1 || c < 1 || out_size < 1 || ksize < 1) {
    set_error("resample_u8: bad arguments"); return -1;
  }
  const uint8_t* s = reinterpret_cast<const uint8_t*>(src);
  uint8_t* d = reinterpret_cast<uint8_t*>(dst);
  if (vertical) {
    launch_k(resample_u8_v_kernel, dim3(grid_1d((size_t)n * out_size * w_in * c)), dim3(256), 0, as_stream(stream), s, d, n, h_in, out_size,
             w_in, c, bounds, kk, ksize);
    return finish_launch("resample_u8 (vertical)");
  }
  launch_k(resample_u8_h_kernel, dim3(grid_1d((size_t)n * h_in * out_size * c)), dim3(256), 0, as_stream(stream), s, d, n, h_in, w_in, out_size,
           c, bounds, kk, ksize);
  return finish_launch("resample_u8 (horizontal)");
}

int clslam_u8_to_f32_chw(const void* hwc, float* chw, int n, int h, int w, void* stream) {
  if (!hwc || !chw || n < 1 || h < 1 || w < 1) { set_error("u8_to_f32_chw: bad arguments"); return -1; }
  launch_k(u8_to_f32_chw_kernel, dim3(grid_1d((size_t)n * h * w)), dim3(256), 0, as_stream(stream), reinterpret_cast<const uint8_t*>(hwc), chw,
           n, h * w);
  return finish_launch("u8_to_f32_chw");
}

long long clslam_color_jitter_workspace_bytes(int n) { return (long long)n * kGrayBlocks * 8; }

int clslam_color_jitter_op(const float* src, float* dst, int n, int h, int w, int op, double factor, void* workspace,
                           long long workspace_bytes, void* stream) {
  if (!src || !dst || n < 1 || h < 1 || w < 1 || op < 0 || op > 3) { set_error("color_jitter_op: bad arguments"); return -1; }
  const int hw = h * w;
  double* part = reinterpret_cast<double*>(workspace);
  if (op == 1) {
    if (!workspace || workspace_bytes < clslam_color_jitter_workspace_bytes(n)) { set_error("color_jitter_op: workspace too small"); return -2; }
    launch_k(gray_partial_kernel, dim3(kGrayBlocks, n), dim3(256), 0, as_stream(stream), src, hw, part);
    if (int rc = finish_launch("gray_partial")) return rc;
  }
  // _blend(img1, img2, ratio): ratio and (1.0 - ratio) are Python floats (doubles) multiplied into fp32 tensors
  const float ratio = (float)factor, inv = (float)(1.0 - factor);
  launch_k(jitter_op_kernel, dim3(grid_1d((size_t)n * hw)), dim3(256), 0, as_stream(stream), src, dst, n, hw, op, ratio, inv,
           (const double*)part, (int)kGrayBlocks);
  return finish_launch("color_jitter_op");
}

}  // extern "C"
