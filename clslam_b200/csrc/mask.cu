// mask_dynamic: pre-training with dynamic-object masks (depth_pose_prediction.py:1062-1068, :1080-1082, :1148-1176; SURVEY.md
// section 8(f) row 3).  Three small kernels next to the photometric / loss kernels (include/clslam_b200.h has the contract):
//   mask_prepare     reprojection weights H*W / (static pixels of the batch), positions of the first b selected gradients per scale
//   smooth_masked    the b smoothness terms per scale AS THE REFERENCE EXECUTES THEM (the i-th selected x- and y-gradient of the
//                    whole batch goes to "sample i"), their loss contributions and their gradient (sparse entries + uniform term)
//   smooth_masked_apply   adds that gradient, times sigmoid', to d(loss)/d(pre-sigmoid disparity)
// All of it is O(b) work plus block reductions; one CTA of 1024 threads, deterministic.
#include "common.cuh"

namespace clslam {

struct Ptr4 { const float* p[4]; };

__device__ __forceinline__ float block_sum_1024(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) for (int i = 0; i < 32; ++i) t += red[i];          // fixed order
  __syncthreads();
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  t = red[0];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(1024)
mask_prepare_kernel(Ptr4 masks, int b, int h, int w, int n_scales, float* __restrict__ reproj_w, int* __restrict__ pos) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  // static pixels of the whole batch at full resolution (exact in fp32: counts < 2^24 per thread-strided partial sum)
  const size_t n0 = (size_t)b * h * w;
  float cnt = 0.f;
  for (size_t i = threadIdx.x; i < n0; i += 1024) cnt += masks.p[0][i] >= 0.5f ? 0.f : 1.f;
  cnt = block_sum_1024(cnt, red);
  if ((int)threadIdx.x < b) reproj_w[threadIdx.x] = (float)h * (float)w / cnt;          // cnt == 0 -> inf -> NaN loss, like the reference's empty mean
  // the first b selected gradients per scale and direction, raster order over the whole batch: one warp per (scale, direction)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 2 * n_scales) {
    const int s = warp >> 1, dir = warp & 1;
    const int hs = h >> s, ws = w >> s;
    const int gh = dir == 0 ? hs : hs - 1, gw = dir == 0 ? ws - 1 : ws;          // x-gradient grid (hs, ws-1), y-gradient grid (hs-1, ws)
    const long long total = (long long)b * gh * gw;
    int found = 0;
    int* out = pos + ((size_t)s * 2 + dir) * b;
    for (long long base = 0; base < total && found < b; base += 32) {
      const long long i = base + lane;
      bool sel = false;
      if (i < total) {
        const int x = (int)(i % gw), y = (int)((i / gw) % gh), bb = (int)(i / ((long long)gw * gh));
        sel = masks.p[s][((size_t)bb * hs + y) * ws + x] < 0.5f;          // mask[..., :-1] / mask[..., :-1, :]: the pixel itself
      }
      const unsigned m = __ballot_sync(0xffffffffu, sel);
      if (sel) {
        const int rank = found + __popc(m & ((1u << lane) - 1u));
        if (rank < b) out[rank] = (int)i;
      }
      found += __popc(m);
    }
    if (lane == 0) for (int k = found; k < b; ++k) out[k] = -1;
  }
}

__global__ void __launch_bounds__(1024)
smooth_masked_kernel(Ptr4 disp, Ptr4 color, const int* __restrict__ pos, const float* __restrict__ sample_w, int b, int h, int w,
                     int n_scales, float disparity_smoothness, float* __restrict__ losses, int* __restrict__ sm_idx,
                     float* __restrict__ sm_val, float* __restrict__ sm_uniform) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  __shared__ float s_den[4][64];          // mean disparity + 1e-7 per (scale, sample); b <= 64
  for (int s = 0; s < n_scales; ++s) {
    const int n = (h >> s) * (w >> s);
    for (int bb = 0; bb < b; ++bb) {
      float t = 0.f;
      const float* d = disp.p[s] + (size_t)bb * n;
      for (int i = threadIdx.x; i < n; i += 1024) t += d[i];
      t = block_sum_1024(t, red);
      if (threadIdx.x == 0) s_den[s][bb] = t / (float)n + 1e-7f;
    }
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  float total_add = 0.f;
  for (int s = 0; s < n_scales; ++s) {
    const int hs = h >> s, ws = w >> s, n = hs * ws;
    const float cscale = (disparity_smoothness / (float)(1 << s)) / (float)n_scales;
    int* idx = sm_idx + (size_t)s * 4 * b;
    float* val = sm_val + (size_t)s * 4 * b;
    float* uni = sm_uniform + (size_t)s * b;
    for (int i = 0; i < 4 * b; ++i) { idx[i] = -1; val[i] = 0.f; }
    for (int i = 0; i < b; ++i) uni[i] = 0.f;
    float smooth = 0.f;
    for (int i = 0; i < b; ++i) {
      const float wgt = sample_w[i];
      for (int dir = 0; dir < 2; ++dir) {
        const int p = pos[((size_t)s * 2 + dir) * b + i];
        if (p < 0) { smooth += __int_as_float(0x7fc00000); continue; }          // the reference's mean of an empty selection
        const int gh = dir == 0 ? hs : hs - 1, gw = dir == 0 ? ws - 1 : ws;
        const int x = p % gw, y = (p / gw) % gh, bb = p / (gw * gh);
        const int q0 = y * ws + x, q1 = dir == 0 ? q0 + 1 : q0 + ws;
        const float* d = disp.p[s] + (size_t)bb * n;
        const float* img = color.p[s] + (size_t)bb * 3 * n;
        const float den = s_den[s][bb];
        float e = 0.f;
        for (int c = 0; c < 3; ++c) e += fabsf(img[(size_t)c * n + q0] - img[(size_t)c * n + q1]);
        e = expf(-(e / 3.f));
        const float dv = d[q0] / den - d[q1] / den;
        smooth += fabsf(dv) * e * wgt;
        const float sg = dv > 0.f ? 1.f : (dv < 0.f ? -1.f : 0.f);
        const float g = cscale * wgt * sg * e;          // d(loss) / d(normalised disparity at q0); -g at q1
        idx[4 * i + 2 * dir] = bb * n + q0;     val[4 * i + 2 * dir] = g / den;
        idx[4 * i + 2 * dir + 1] = bb * n + q1; val[4 * i + 2 * dir + 1] = -g / den;
        // through the mean: nd = d / den, d(den) / d(d[q]) = 1 / n for every pixel of the sample
        uni[bb] -= (g * d[q0] - g * d[q1]) / (den * den) / (float)n;
      }
    }
    const float reg = disparity_smoothness / (float)(1 << s) * smooth;
    losses[4 + s] = smooth;
    losses[8 + s] = reg;
    losses[12 + s] += reg;
    total_add += reg / (float)n_scales;
  }
  losses[16] += total_add;
  losses[18] += total_add;
  losses[19] = (losses[18] != losses[18]) ? 1.f : 0.f;
}

__global__ void __launch_bounds__(256)
smooth_masked_apply_kernel(const float* __restrict__ disp, const int* __restrict__ sm_idx, const float* __restrict__ sm_val,
                           const float* __restrict__ sm_uniform, int b, int n, float* __restrict__ dz) {
  pdl_wait();
  pdl_trigger();
  __shared__ int s_idx[256];
  __shared__ float s_val[256];
  const int ne = 4 * b;                    // <= 256 (b <= 64)
  if ((int)threadIdx.x < ne) { s_idx[threadIdx.x] = sm_idx[threadIdx.x]; s_val[threadIdx.x] = sm_val[threadIdx.x]; }
  __syncthreads();
  const int total = b * n;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
    float g = sm_uniform[q / n];
    for (int e = 0; e < ne; ++e) g += s_idx[e] == q ? s_val[e] : 0.f;          // fixed order: deterministic
    const float dv = disp[q];
    dz[q] += g * dv * (1.f - dv);
  }
}

}  // namespace clslam

using namespace clslam;

extern "C" {

int clslam_mask_prepare(const float* const* masks, int b, int h, int w, int n_scales, float* reproj_w, int* pos, void* stream) {
  if (!masks || !reproj_w || !pos || b < 1 || b > 64 || n_scales < 1 || n_scales > 4) { set_error("mask_prepare: bad arguments (b <= 64)"); return -1; }
  Ptr4 m{};
  for (int s = 0; s < n_scales; ++s) { if (!masks[s]) { set_error("mask_prepare: null mask"); return -1; } m.p[s] = masks[s]; }
  launch_k(mask_prepare_kernel, dim3(1), dim3(1024), 0, as_stream(stream), m, b, h, w, n_scales, reproj_w, pos);
  return finish_launch("mask_prepare");
}

int clslam_smooth_masked(const float* const* disp, const float* const* color, const int* pos, const float* sample_w, int b, int h,
                         int w, int n_scales, float disparity_smoothness, float* losses, int* sm_idx, float* sm_val,
                         float* sm_uniform, void* stream) {
  if (!disp || !color || !pos || !sample_w || !losses || !sm_idx || !sm_val || !sm_uniform || b < 1 || b > 64 || n_scales < 1 ||
      n_scales > 4) { set_error("smooth_masked: bad arguments (b <= 64)"); return -1; }
  Ptr4 d{}, c{};
  for (int s = 0; s < n_scales; ++s) { d.p[s] = disp[s]; c.p[s] = color[s]; }
  launch_k(smooth_masked_kernel, dim3(1), dim3(1024), 0, as_stream(stream), d, c, pos, sample_w, b, h, w, n_scales,
           disparity_smoothness, losses, sm_idx, sm_val, sm_uniform);
  return finish_launch("smooth_masked");
}

int clslam_smooth_masked_apply(const float* disp, const int* sm_idx, const float* sm_val, const float* sm_uniform, int b, int h,
                               int w, int scale, float* dz_disp, void* stream) {
  if (b < 1 || b > 64) { set_error("smooth_masked_apply: b <= 64"); return -1; }
  const int n = (h >> scale) * (w >> scale);
  int blocks = (b * n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(smooth_masked_apply_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), disp, sm_idx + (size_t)scale * 4 * b,
           sm_val + (size_t)scale * 4 * b, sm_uniform + (size_t)scale * b, b, n, dz_disp);
  return finish_launch("smooth_masked_apply");
}

}  // extern "C"
