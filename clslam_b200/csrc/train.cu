// Kernels of the full-network training path: what `train()` /
// `_set_adapt(freeze_encoder=False)` need on top of the adaptation step -- BatchNorm2d with batch statistics (forward,
// running-stat update, backward), the backward of the 3x3 / stride-2 max pool, and gradient accumulation.  All memory-bound,
// fp32 NHWC, deterministic (per-block partial sums in double, fixed-order final sums).
// Reference: torchvision ResNet-18 BasicBlock / nn.BatchNorm2d (eps 1e-5, momentum 0.1) as driven by
// depth_pose_prediction.py:219-289 (train), :797-800 (_set_train), :807-827 (_set_adapt).
#include "common.cuh"

namespace clslam {

constexpr int kBnBlocks = 296;      // partial-sum blocks per statistics pass (2 per SM)

// ---- per-channel sums over the pixels of each group -----------------------------------------------------------------------
// x: (n, hw, c) fp32 NHWC; the n images are split into `groups` equal groups with separate statistics (the pose encoder runs
// its two passes of B images as one batch of 2B: nn.BatchNorm2d sees them one pass at a time, depth_pose_prediction.py:956-963)
// MODE 0: sum x, sum x^2.  MODE 1 (backward): g = dy * act'(y); sum g, sum g * xhat with xhat = (x - mean) * invstd.
// part: [groups][blocks][2][c] doubles
template <int MODE>
__global__ void __launch_bounds__(256)
bn_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y, int relu,
                  const float* __restrict__ mean, const float* __restrict__ invstd, long long px_per_group, int c,
                  double* __restrict__ part) {
  pdl_wait();
  pdl_trigger();
  const int quads = c / 4, lanes = 256 / quads;
  const int qd = threadIdx.x % quads, pl = threadIdx.x / quads;
  const int grp = blockIdx.y;
  const long long per = (px_per_group + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per, p1 = p0 + per < px_per_group ? p0 + per : px_per_group;
  const long long base = (long long)grp * px_per_group;
  double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  float4 mu = f4zero(), is = f4zero();
  if (MODE == 1) { mu = ld4(mean + grp * c + qd * 4); is = ld4(invstd + grp * c + qd * 4); }
  if (pl < lanes) {
    for (long long p = p0 + pl; p < p1; p += lanes) {
      const size_t off = (size_t)(base + p) * c + qd * 4;
      const float4 v = ld4(x + off);
      const float xv[4] = {v.x, v.y, v.z, v.w};
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { s0[j] += (double)xv[j]; s1[j] += (double)xv[j] * (double)xv[j]; }
      } else {
        const float4 g4 = ld4(dy + off);
        float gv[4] = {g4.x, g4.y, g4.z, g4.w};
        if (relu) {
          const float4 y4 = ld4(y + off);
          const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) gv[j] = yv[j] > 0.f ? gv[j] : 0.f;
        }
        const float mv[4] = {mu.x, mu.y, mu.z, mu.w}, iv[4] = {is.x, is.y, is.z, is.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xh = (xv[j] - mv[j]) * iv[j];
          s0[j] += (double)gv[j];
          s1[j] += (double)gv[j] * (double)xh;
        }
      }
    }
  }
  __shared__ double red[2][256][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { red[0][threadIdx.x][j] = s0[j]; red[1][threadIdx.x][j] = s1[j]; }
  __syncthreads();
  if ((int)threadIdx.x < quads) {
    double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
    for (int k = 0; k < lanes; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) { a[j] += red[0][k * quads + threadIdx.x][j]; b[j] += red[1][k * quads + threadIdx.x][j]; }
    double* o = part + (((size_t)grp * gridDim.x + blockIdx.x) * 2) * c + threadIdx.x * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[j] = a[j]; o[c + j] = b[j]; }
  }
}

// forward statistics: mean / invstd per (group, channel) from the partial sums; training mode also updates the running
// statistics once per group, in group order (momentum 0.1, unbiased variance), like two consecutive module calls
__global__ void __launch_bounds__(256)
bn_stats_final_kernel(const double* __restrict__ part, int blocks, int groups, long long px_per_group, int c, float eps,
                      float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                      float* __restrict__ mean, float* __restrict__ invstd) {
  pdl_wait();
  pdl_trigger();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  for (int grp = 0; grp < groups; ++grp) {
    double s = 0.0, ss = 0.0;
    for (int b = 0; b < blocks; ++b) {
      const double* o = part + (((size_t)grp * blocks + b) * 2) * c;
      s += o[ch]; ss += o[c + ch];
    }
    const double n = (double)px_per_group;
    const double m = s / n;
    double var = ss / n - m * m;
    if (var < 0.0) var = 0.0;
    mean[grp * c + ch] = (float)m;
    invstd[grp * c + ch] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
      running_mean[ch] = (float)((1.0 - (double)momentum) * (double)running_mean[ch] + (double)momentum * m);
      running_var[ch] = (float)((1.0 - (double)momentum) * (double)running_var[ch] + (double)momentum * unbiased);
    }
  }
}

// evaluation mode: mean / invstd from the running statistics (every group the same)
__global__ void __launch_bounds__(256)
bn_stats_eval_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var, int groups, int c, float eps,
                     float* __restrict__ mean, float* __restrict__ invstd) {
  pdl_wait();
  pdl_trigger();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  for (int grp = 0; grp < groups; ++grp) {
    mean[grp * c + ch] = running_mean[ch];
    invstd[grp * c + ch] = 1.f / sqrtf(running_var[ch] + eps);
  }
}

// y = act((x - mean) * invstd * gamma + beta (+ residual))
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ residual, int relu,
                long long px_per_group, int groups, int c, float* __restrict__ y) {
  pdl_wait();
  pdl_trigger();
  const int c4 = c / 4;
  const size_t total = (size_t)groups * px_per_group * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c4) * 4;
    const int grp = (int)((i / c4) / px_per_group);
    const float4 v = ld4(x + i * 4), mu = ld4(mean + grp * c + cc), is = ld4(invstd + grp * c + cc);
    const float4 ga = ld4(gamma + cc), be = ld4(beta + cc);
    float4 o;
    o.x = (v.x - mu.x) * is.x * ga.x + be.x; o.y = (v.y - mu.y) * is.y * ga.y + be.y;
    o.z = (v.z - mu.z) * is.z * ga.z + be.z; o.w = (v.w - mu.w) * is.w * ga.w + be.w;
    if (residual) { const float4 r = ld4(residual + i * 4); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    *reinterpret_cast<float4*>(y + i * 4) = o;
  }
}

// backward sums -> d(gamma), d(beta) (summed over the groups: one parameter, several calls) and the per-group means the input
// gradient needs
__global__ void __launch_bounds__(256)
bn_bwd_final_kernel(const double* __restrict__ part, int blocks, int groups, long long px_per_group, int c,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ mean_g, float* __restrict__ mean_gx) {
  pdl_wait();
  pdl_trigger();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  double tg = 0.0, tgx = 0.0;
  for (int grp = 0; grp < groups; ++grp) {
    double s = 0.0, ss = 0.0;
    for (int b = 0; b < blocks; ++b) {
      const double* o = part + (((size_t)grp * blocks + b) * 2) * c;
      s += o[ch]; ss += o[c + ch];
    }
    mean_g[grp * c + ch] = (float)(s / (double)px_per_group);
    mean_gx[grp * c + ch] = (float)(ss / (double)px_per_group);
    tg += s; tgx += ss;
  }
  if (dbeta) dbeta[ch] = (float)tg;
  if (dgamma) dgamma[ch] = (float)tgx;
}

// dx = gamma * invstd * (g - mean_g - xhat * mean_gx)   (training mode; batch_stats = 0: gamma * invstd * g), g = dy * act'(y)
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y, int relu,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ mean_g, const float* __restrict__ mean_gx, int batch_stats, long long px_per_group,
                    int groups, int c, float* __restrict__ dx, float* __restrict__ g_out) {
  pdl_wait();
  pdl_trigger();
  const int c4 = c / 4;
  const size_t total = (size_t)groups * px_per_group * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c4) * 4;
    const int grp = (int)((i / c4) / px_per_group);
    const float4 x4 = ld4(x + i * 4), g4 = ld4(dy + i * 4);
    float xv[4] = {x4.x, x4.y, x4.z, x4.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w};
    if (relu) {
      const float4 y4 = ld4(y + i * 4);
      const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) gv[j] = yv[j] > 0.f ? gv[j] : 0.f;
    }
    if (g_out) *reinterpret_cast<float4*>(g_out + i * 4) = make_float4(gv[0], gv[1], gv[2], gv[3]);      // the identity branch of a BasicBlock
    const float4 mu = ld4(mean + grp * c + cc), is = ld4(invstd + grp * c + cc), ga = ld4(gamma + cc);
    const float mv[4] = {mu.x, mu.y, mu.z, mu.w}, iv[4] = {is.x, is.y, is.z, is.w}, gav[4] = {ga.x, ga.y, ga.z, ga.w};
    float o[4];
    if (batch_stats) {
      const float4 a = ld4(mean_g + grp * c + cc), b = ld4(mean_gx + grp * c + cc);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = (xv[j] - mv[j]) * iv[j];
        o[j] = gav[j] * iv[j] * (gv[j] - av[j] - xh * bv[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = gav[j] * iv[j] * gv[j];
    }
    *reinterpret_cast<float4*>(dx + i * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// backward of MaxPool2d(3, stride 2, padding 1), NHWC: every input pixel collects dy of the (up to four) windows whose maximum
// it is.  Ties go to the first maximum in row-major window order, as in ATen's max_pool2d_with_indices.
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int n, int hi, int wi, int c, int ho, int wo,
                   float* __restrict__ dx) {
  pdl_wait();
  pdl_trigger();
  const int c4 = c / 4;
  const size_t total = (size_t)n * hi * wi * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c4) * 4;
    size_t p = i / c4;
    const int iw = (int)(p % wi); p /= wi;
    const int ih = (int)(p % hi);
    const int b = (int)(p / hi);
    const float4 v4 = ld4(x + i * 4);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // windows (oh, ow) that contain (ih, iw): 2*oh - 1 <= ih <= 2*oh + 1
    for (int oh = (ih) / 2; oh <= (ih + 1) / 2; ++oh) {
      if (oh < 0 || oh >= ho) continue;
      for (int ow = (iw) / 2; ow <= (iw + 1) / 2; ++ow) {
        if (ow < 0 || ow >= wo) continue;
        // is (ih, iw) the FIRST maximum of window (oh, ow)?
        bool first[4] = {true, true, true, true};
        for (int r = 0; r < 3; ++r) {
          const int yy = 2 * oh + r - 1;
          if (yy < 0 || yy >= hi) continue;
          for (int s = 0; s < 3; ++s) {
            const int xx = 2 * ow + s - 1;
            if (xx < 0 || xx >= wi || (yy == ih && xx == iw)) continue;
            const float4 q4 = ld4(x + (((size_t)b * hi + yy) * wi + xx) * c + cc);
            const float q[4] = {q4.x, q4.y, q4.z, q4.w};
            const bool before = yy < ih || (yy == ih && xx < iw);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (q[j] > v[j] || (before && q[j] == v[j])) first[j] = false;
          }
        }
        const float4 g4 = ld4(dy + (((size_t)b * ho + oh) * wo + ow) * c + cc);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) if (first[j]) acc[j] += g[j];
      }
    }
    *reinterpret_cast<float4*>(dx + i * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// out = a + b (+ c), n elements (multiples of 4): gradient accumulation at the joins of the ResNet / U-Net graph
__global__ void __launch_bounds__(256)
add3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c3, float* __restrict__ out, long long n4) {
  pdl_wait();
  pdl_trigger();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = ld4(a + i * 4);
    if (b) f4add(v, ld4(b + i * 4));
    if (c3) f4add(v, ld4(c3 + i * 4));
    *reinterpret_cast<float4*>(out + i * 4) = v;
  }
}

static int grid_for_tr(size_t work, int cap = 148 * 8) {
  size_t b = (work + 255) / 256;
  if (b < 1) b = 1;
  if (b > (size_t)cap) b = cap;
  return (int)b;
}

static int bn_check(long long px_per_group, int groups, int c, const char* what) {
  if (px_per_group < 1 || groups < 1 || groups > 4 || c < 4 || c % 4 || c > 1024 || (256 % (c / 4) != 0 && c / 4 < 256)) {
    set_error("%s: unsupported shape (pixels per group %lld, groups %d, channels %d)", what, px_per_group, groups, c); return -1;
  }
  return 0;
}

}  // namespace clslam

using namespace clslam;

extern "C" {

long long clslam_bn_workspace_bytes(int groups, int c) {
  // partial sums [groups][kBnBlocks][2][c] doubles + mean_g / mean_gx [2][groups][c] floats
  return (long long)groups * kBnBlocks * 2 * c * 8 + 2LL * groups * c * 4;
}

int clslam_bn_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var, int training,
                  float momentum, float eps, const float* residual, int relu, int n, int hw, int c, int groups, float* y,
                  float* save_mean, float* save_invstd, void* workspace, long long workspace_bytes, void* stream) {
  if (!x || !gamma || !beta || !running_mean || !running_var || !y || !save_mean || !save_invstd || n % groups) {
    set_error("bn_fwd: bad arguments"); return -1;
  }
  const long long ppg = (long long)(n / groups) * hw;
  if (int rc = bn_check(ppg, groups, c, "bn_fwd")) return rc;
  cudaStream_t st = as_stream(stream);
  if (training) {
    if (!workspace || workspace_bytes < clslam_bn_workspace_bytes(groups, c)) { set_error("bn_fwd: workspace too small"); return -2; }
    double* part = reinterpret_cast<double*>(workspace);
    launch_k(bn_partial_kernel<0>, dim3(kBnBlocks, groups), dim3(256), 0, st, x, (const float*)nullptr, (const float*)nullptr, 0,
             (const float*)nullptr, (const float*)nullptr, ppg, c, part);
    if (int rc = finish_launch("bn_fwd (partial sums)")) return rc;
    launch_k(bn_stats_final_kernel, dim3((c + 255) / 256), dim3(256), 0, st, (const double*)part, (int)kBnBlocks, groups, ppg, c, eps, momentum,
             running_mean, running_var, save_mean, save_invstd);
    if (int rc = finish_launch("bn_fwd (statistics)")) return rc;
  } else {
    launch_k(bn_stats_eval_kernel, dim3((c + 255) / 256), dim3(256), 0, st, (const float*)running_mean, (const float*)running_var, groups, c, eps,
             save_mean, save_invstd);
    if (int rc = finish_launch("bn_fwd (running statistics)")) return rc;
  }
  launch_k(bn_apply_kernel, dim3(grid_for_tr((size_t)groups * ppg * (c / 4))), dim3(256), 0, st, x, (const float*)save_mean,
           (const float*)save_invstd, gamma, beta, residual, relu, ppg, groups, c, y);
  return finish_launch("bn_fwd (apply)");
}

int clslam_bn_bwd(const float* x, const float* dy, const float* y, int relu, const float* gamma, const float* save_mean,
                  const float* save_invstd, int training, int n, int hw, int c, int groups, float* dx, float* dgamma, float* dbeta,
                  float* g_out, void* workspace, long long workspace_bytes, void* stream) {
  if (!x || !dy || !gamma || !save_mean || !save_invstd || !dx || n % groups || (relu && !y)) { set_error("bn_bwd: bad arguments"); return -1; }
  const long long ppg = (long long)(n / groups) * hw;
  if (int rc = bn_check(ppg, groups, c, "bn_bwd")) return rc;
  if (!workspace || workspace_bytes < clslam_bn_workspace_bytes(groups, c)) { set_error("bn_bwd: workspace too small"); return -2; }
  cudaStream_t st = as_stream(stream);
  double* part = reinterpret_cast<double*>(workspace);
  float* mean_g = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + (size_t)groups * kBnBlocks * 2 * c * 8);
  float* mean_gx = mean_g + (size_t)groups * c;
  launch_k(bn_partial_kernel<1>, dim3(kBnBlocks, groups), dim3(256), 0, st, x, dy, y, relu, save_mean, save_invstd, ppg, c, part);
  if (int rc = finish_launch("bn_bwd (partial sums)")) return rc;
  launch_k(bn_bwd_final_kernel, dim3((c + 255) / 256), dim3(256), 0, st, (const double*)part, (int)kBnBlocks, groups, ppg, c, dgamma, dbeta,
           mean_g, mean_gx);
  if (int rc = finish_launch("bn_bwd (sums)")) return rc;
  launch_k(bn_bwd_apply_kernel, dim3(grid_for_tr((size_t)groups * ppg * (c / 4))), dim3(256), 0, st, x, dy, y, relu, save_mean, save_invstd,
           gamma, (const float*)mean_g, (const float*)mean_gx, training, ppg, groups, c, dx, g_out);
  return finish_launch("bn_bwd (apply)");
}

int clslam_maxpool3x3s2_bwd(const float* x, const float* dy, float* dx, int n, int hi, int wi, int c, void* stream) {
  if (!x || !dy || !dx || c % 4) { set_error("maxpool_bwd: bad arguments"); return -1; }
  const int ho = (hi + 2 - 3) / 2 + 1, wo = (wi + 2 - 3) / 2 + 1;
  launch_k(maxpool_bwd_kernel, dim3(grid_for_tr((size_t)n * hi * wi * (c / 4))), dim3(256), 0, as_stream(stream), x, dy, n, hi, wi, c, ho, wo, dx);
  return finish_launch("maxpool_bwd");
}

int clslam_add3(const float* a, const float* b, const float* c, float* out, long long n, void* stream) {
  if (!a || !out || n < 4 || n % 4) { set_error("add3: bad arguments"); return -1; }
  launch_k(add3_kernel, dim3(grid_for_tr((size_t)(n / 4))), dim3(256), 0, as_stream(stream), a, b, c, out, n / 4);
  return finish_launch("add3");
}

}  // extern "C"
