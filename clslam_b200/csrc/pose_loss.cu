// Tiny O(batch) kernels around the pose head and the loss assembly:
//   pose_head_fwd / pose_head_bwd : spatial mean * 0.01 -> Rodrigues/translation -> cam_T_cam -> K.T
//   loss_finalize                 : 19 loss values incl. the smoothness term as the reference executes it
// References: networks/pose_decoder.py:49-54, utils.py:34-117, networks/layers.py:94,
// depth_pose_prediction.py:1071-1146.
#include "common.cuh"

namespace clslam {

struct Se3Fwd {
  float theta, inv, x, y, z, ca, sa, C, xC, yC, zC;
  float R[9];
};

__device__ void se3_forward(const float* aa, Se3Fwd& s) {
  s.theta = sqrtf(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);
  const float den = s.theta + 1e-7f;
  s.inv = 1.f / den;
  s.x = aa[0] / den; s.y = aa[1] / den; s.z = aa[2] / den;
  s.ca = cosf(s.theta); s.sa = sinf(s.theta);
  s.C = 1.f - s.ca;
  const float xs = s.x * s.sa, ys = s.y * s.sa, zs = s.z * s.sa;
  s.xC = s.x * s.C; s.yC = s.y * s.C; s.zC = s.z * s.C;
  const float xyC = s.x * s.yC, yzC = s.y * s.zC, zxC = s.z * s.xC;
  s.R[0] = s.x * s.xC + s.ca; s.R[1] = xyC - zs;          s.R[2] = zxC + ys;
  s.R[3] = xyC + zs;          s.R[4] = s.y * s.yC + s.ca; s.R[5] = yzC - xs;
  s.R[6] = zxC - ys;          s.R[7] = yzC + xs;          s.R[8] = s.z * s.zC + s.ca;
}

// cam_T_cam (row-major 4x4) from R (3x3), t, invert flag (utils.py:34-55)
__device__ void se3_compose(const float* R, const float* t, int invert, float* M) {
#pragma unroll
  for (int i = 0; i < 16; ++i) M[i] = 0.f;
  M[15] = 1.f;
  if (!invert) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) M[i * 4 + j] = R[i * 3 + j];
      M[i * 4 + 3] = t[i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        M[i * 4 + k] = R[k * 3 + i];
        acc += R[k * 3 + i] * (-t[k]);
      }
      M[i * 4 + 3] = acc;
    }
  }
}

__global__ void __launch_bounds__(32)
pose_head_fwd_kernel(const float* __restrict__ pose_out, const float* __restrict__ k_mat, int b, int h, int w,
                     int invert_mask, float* __restrict__ axis_angle, float* __restrict__ translation,
                     float* __restrict__ cam_t_cam, float* __restrict__ proj) {
  pdl_wait();
  pdl_trigger();
  const int pass = blockIdx.x / b, bb = blockIdx.x % b;
  __shared__ float out6[6];
  const int lane = threadIdx.x;
  if (lane < 6) {
    const float* p = pose_out + (size_t)blockIdx.x * h * w * 12 + lane;
    float acc_h = 0.f;
    for (int y = 0; y < h; ++y) {           // out.mean(3).mean(2): over W, then over H
      float acc_w = 0.f;
      for (int x = 0; x < w; ++x) acc_w += p[((size_t)y * w + x) * 12];
      acc_h += acc_w / (float)w;
    }
    out6[lane] = 0.01f * (acc_h / (float)h);
  }
  __syncwarp();
  if (lane == 0) {
    const size_t o = (size_t)pass * b + bb;
    float aa[3] = {out6[0], out6[1], out6[2]}, t[3] = {out6[3], out6[4], out6[5]};
    for (int i = 0; i < 3; ++i) { axis_angle[o * 3 + i] = aa[i]; translation[o * 3 + i] = t[i]; }
    Se3Fwd s;
    se3_forward(aa, s);
    float M[16];
    se3_compose(s.R, t, (invert_mask >> pass) & 1, M);
    for (int i = 0; i < 16; ++i) cam_t_cam[o * 16 + i] = M[i];
    const float* K = k_mat + (size_t)bb * 16;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) {
        float acc = 0.f;
        for (int k = 0; k < 4; ++k) acc += K[r * 4 + k] * M[k * 4 + c];
        proj[o * 12 + r * 4 + c] = acc;
      }
  }
}

__global__ void __launch_bounds__(256)
pose_head_bwd_kernel(const float* __restrict__ dproj_partials, int n_partials, const float* __restrict__ k_mat,
                     const float* __restrict__ axis_angle, const float* __restrict__ translation,
                     const double* __restrict__ rel_dist, const float* __restrict__ sample_w, float velocity_scale,
                     int b, int h, int w, int invert_mask, float* __restrict__ dz) {
  pdl_wait();
  pdl_trigger();
  const int pass = blockIdx.x / b, bb = blockIdx.x % b;
  const int lane = threadIdx.x;
  __shared__ float dP[12];
  __shared__ float dout[12];
  __shared__ float part[21][12];
  if (lane < 252) {                       // 21 groups x 12 components, fixed summation order
    const int comp = lane % 12, grp = lane / 12;
    float acc = 0.f;
    for (int p = grp; p < n_partials; p += 21) acc += dproj_partials[(((size_t)p * 2 + pass) * b + bb) * 12 + comp];
    part[grp][comp] = acc;
  }
  __syncthreads();
  if (lane < 12) {
    float acc = 0.f;
    for (int g = 0; g < 21; ++g) acc += part[g][lane];
    dP[lane] = acc;
    dout[lane] = 0.f;
  }
  __syncthreads();
  if (lane == 0) {
    const size_t o = (size_t)pass * b + bb;
    const float* K = k_mat + (size_t)bb * 16;
    const float aa[3] = {axis_angle[o * 3], axis_angle[o * 3 + 1], axis_angle[o * 3 + 2]};
    const float t[3] = {translation[o * 3], translation[o * 3 + 1], translation[o * 3 + 2]};
    const int invert = (invert_mask >> pass) & 1;
    // dM[k][c] = sum_r K[r][k] * dP[r][c], rows k < 3 matter (row 3 of M is constant)
    float dM[12];
    for (int k = 0; k < 3; ++k)
      for (int c = 0; c < 4; ++c) {
        float acc = 0.f;
        for (int r = 0; r < 3; ++r) acc += K[r * 4 + k] * dP[r * 4 + c];
        dM[k * 4 + c] = acc;
      }
    Se3Fwd s;
    se3_forward(aa, s);
    float dR[9], dt[3];
    if (!invert) {
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) dR[i * 3 + j] = dM[i * 4 + j];
        dt[i] = dM[i * 4 + 3];
      }
    } else {
      // M[i][j] = R[j][i];  M[i][3] = sum_k R[k][i] * (-t[k])
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) dR[j * 3 + i] = dM[i * 4 + j] + dM[i * 4 + 3] * (-t[j]);
      for (int k = 0; k < 3; ++k) {
        float acc = 0.f;
        for (int i = 0; i < 3; ++i) acc += dM[i * 4 + 3] * s.R[k * 3 + i];
        dt[k] = -acc;
      }
    }
    // velocity term: w_b * scale * 0.5 * | ||t|| - |gt| |   (pass 0 <-> relative_distance 0, pass 1 <-> 1)
    if (velocity_scale > 0.f && rel_dist) {
      const float nrm = sqrtf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
      const double diff = (double)nrm - fabs(rel_dist[(size_t)pass * b + bb]);
      const float sg = diff > 0.0 ? 1.f : (diff < 0.0 ? -1.f : 0.f);
      if (nrm > 0.f) {
        const float coef = sample_w[bb] * velocity_scale * 0.5f * sg / nrm;
        for (int i = 0; i < 3; ++i) dt[i] += coef * t[i];
      }
    }
    // Rodrigues backward
    float dx = 0.f, dy = 0.f, dzv = 0.f, dxC = 0.f, dyC = 0.f, dzC = 0.f, dca = 0.f, dsa = 0.f, dC = 0.f;
    dca = dR[0] + dR[4] + dR[8];
    dx += dR[0] * s.xC; dxC += dR[0] * s.x;
    dy += dR[4] * s.yC; dyC += dR[4] * s.y;
    dzv += dR[8] * s.zC; dzC += dR[8] * s.z;
    const float dxyC = dR[1] + dR[3], dzs = -dR[1] + dR[3];
    const float dzxC = dR[2] + dR[6], dys = dR[2] - dR[6];
    const float dyzC = dR[5] + dR[7], dxs = -dR[5] + dR[7];
    dx += dxyC * s.yC; dyC += dxyC * s.x;
    dy += dyzC * s.zC; dzC += dyzC * s.y;
    dzv += dzxC * s.xC; dxC += dzxC * s.z;
    dx += dxs * s.sa; dsa += dxs * s.x;
    dy += dys * s.sa; dsa += dys * s.y;
    dzv += dzs * s.sa; dsa += dzs * s.z;
    dx += dxC * s.C; dC += dxC * s.x;
    dy += dyC * s.C; dC += dyC * s.y;
    dzv += dzC * s.C; dC += dzC * s.z;
    dca -= dC;
    float dtheta = -s.sa * dca + s.ca * dsa;
    const float dax[3] = {dx, dy, dzv};
    float daa[3];
    float dot = 0.f;
    for (int k = 0; k < 3; ++k) dot += dax[k] * aa[k];
    dtheta += -dot * s.inv * s.inv;
    for (int k = 0; k < 3; ++k) daa[k] = dax[k] * s.inv + (s.theta > 0.f ? dtheta * aa[k] / s.theta : 0.f);
    const float k_mean = 0.01f / ((float)h * (float)w);
    for (int k = 0; k < 3; ++k) { dout[k] = daa[k] * k_mean; dout[3 + k] = dt[k] * k_mean; }
  }
  __syncthreads();
  float* d = dz + (size_t)blockIdx.x * h * w * 12;
  for (int i = lane; i < h * w * 12; i += 256) d[i] = dout[i % 12];
}

// ---- loss assembly ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
loss_finalize_kernel(clslam_loss_args a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  __shared__ float s_mean[4];
  __shared__ float s_reproj[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = a.n_scales;
  const float hw = (float)a.h * (float)a.w;
  // reprojection terms: sum_b w_b * mean_hw(min-loss)
  for (int s = 0; s < S; ++s) {
    float acc = 0.f;
    for (int bb = warp; bb < a.b; bb += 32) {
      float t = 0.f;
      for (int i = lane; i < a.n_tiles; i += 32) t += a.loss_partials[((size_t)s * a.b + bb) * a.n_tiles + i];
      t = warp_sum(t);
      acc += (a.reproj_w ? a.reproj_w : a.sample_w)[bb] * (t / hw);
    }
    __syncthreads();
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int i = 0; i < 32; ++i) t += red[i]; s_reproj[s] = t; }
    // mean disparity of global sample 0 at this scale
    if (a.has_sample0) {
      const int n = (a.h >> s) * (a.w >> s);       // multiple of 4 (h, w are multiples of 32)
      float t = 0.f;
      const float4* d4 = reinterpret_cast<const float4*>(a.disp[s]);
      for (int i = tid; i < n / 4; i += 1024) { const float4 q = d4[i]; t += (q.x + q.y) + (q.z + q.w); }
      t = warp_sum(t);
      __syncthreads();
      if (lane == 0) red[warp] = t;
      __syncthreads();
      if (tid == 0) { float u = 0.f; for (int i = 0; i < 32; ++i) u += red[i]; s_mean[s] = u / (float)n; }
    }
    __syncthreads();
  }
  if (tid != 0) return;
  float total = 0.f;
  const int Bg = a.b_global;
  for (int s = 0; s < S; ++s) {
    const int hs = a.h >> s, ws = a.w >> s;
    float smooth = 0.f;
    float* sg = a.smooth_grad + (size_t)s * (2 * (Bg + 1) + 1);
    for (int i = 0; i < 2 * (Bg + 1) + 1; ++i) sg[i] = 0.f;
    if (a.has_sample0) {
      const float* dsp = a.disp[s];            // sample 0, (hs, ws)
      const float* img = a.color[s];           // sample 0, (3, hs, ws)
      const float den = s_mean[s] + 1e-7f;
      const float cscale = (a.disparity_smoothness / (float)(1 << s)) / (float)S;
      float dm = 0.f;
      for (int i = 0; i < Bg; ++i) {
        const float n00 = dsp[i] / den, n01 = dsp[i + 1] / den, n10 = dsp[ws + i] / den;
        float ex = 0.f, ey = 0.f;
        for (int c = 0; c < 3; ++c) {
          const float* pl = img + (size_t)c * hs * ws;
          ex += fabsf(pl[i] - pl[i + 1]);
          ey += fabsf(pl[i] - pl[ws + i]);
        }
        ex = expf(-(ex / 3.f)); ey = expf(-(ey / 3.f));
        const float dxv = n00 - n01, dyv = n00 - n10;
        const float wgt = a.sample_w_global[i];
        smooth += (fabsf(dxv) * ex + fabsf(dyv) * ey) * wgt;
        const float ci = cscale * wgt;
        const float sx = dxv > 0.f ? 1.f : (dxv < 0.f ? -1.f : 0.f);
        const float sy = dyv > 0.f ? 1.f : (dyv < 0.f ? -1.f : 0.f);
        // gradient wrt the normalised disparity at (0,i), (0,i+1), (1,i)
        sg[i] += ci * (sx * ex + sy * ey);
        sg[i + 1] -= ci * sx * ex;
        sg[(Bg + 1) + i] -= ci * sy * ey;
      }
      // through nd = disp / (mean + eps): direct part / den, mean part uniform over the sample
      for (int i = 0; i <= Bg; ++i) {
        dm -= sg[i] * dsp[i] / (den * den);
        if (i < Bg) dm -= sg[(Bg + 1) + i] * dsp[ws + i] / (den * den);
        sg[i] /= den;
        sg[(Bg + 1) + i] /= den;
      }
      sg[2 * (Bg + 1)] = dm / ((float)hs * (float)ws);
    }
    const float reg = a.disparity_smoothness / (float)(1 << s) * smooth;
    a.losses[0 + s] = s_reproj[s];
    a.losses[4 + s] = smooth;
    a.losses[8 + s] = reg;
    a.losses[12 + s] = s_reproj[s] + reg;
    total += s_reproj[s] + reg;
  }
  total /= (float)S;
  float vel = 0.f;
  if (a.velocity_scale > 0.f && a.rel_dist) {
    for (int bb = 0; bb < a.b; ++bb) {
      float v = 0.f;
      for (int p = 0; p < 2; ++p) {
        const float* t = a.translation + ((size_t)p * a.b + bb) * 3;
        const float nrm = sqrtf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
        v += (float)fabs((double)nrm - fabs(a.rel_dist[(size_t)p * a.b + bb]));
      }
      vel += a.velocity_scale * (v / 2.f) * a.sample_w[bb];
    }
  }
  total += vel;
  a.losses[16] = total;
  a.losses[17] = vel;
  a.losses[18] = total;
  a.losses[19] = (total != total) ? 1.f : 0.f;
}

}  // namespace clslam

using namespace clslam;

extern "C" {

int clslam_pose_head_fwd(const float* pose_out, const float* k_mat, int b, int h, int w, int invert_mask, float* axis_angle,
                         float* translation, float* cam_t_cam, float* proj, void* stream) {
  launch_k(pose_head_fwd_kernel, dim3(2 * b), dim3(32), 0, as_stream(stream), pose_out, k_mat, b, h, w, invert_mask, axis_angle, translation,
                                                            cam_t_cam, proj);
  return finish_launch("pose_head_fwd");
}

int clslam_pose_head_bwd(const float* dproj_partials, int n_partials, const float* k_mat, const float* axis_angle,
                         const float* translation, const double* rel_dist, const float* sample_w, float velocity_scale,
                         int b, int h, int w, int invert_mask, float* dz_pose_out, void* stream) {
  launch_k(pose_head_bwd_kernel, dim3(2 * b), dim3(256), 0, as_stream(stream), dproj_partials, n_partials, k_mat, axis_angle, translation,
                                                            rel_dist, sample_w, velocity_scale, b, h, w, invert_mask,
                                                            dz_pose_out);
  return finish_launch("pose_head_bwd");
}

int clslam_loss_finalize(const clslam_loss_args* a, void* stream) {
  if (!a || a->n_scales < 1 || a->n_scales > 4) { set_error("loss_finalize: bad args"); return -1; }
  if (a->has_sample0 && (a->b_global + 1 > (a->w >> (a->n_scales - 1)) || (a->h >> (a->n_scales - 1)) < 2)) {
    set_error("loss_finalize: reference smoothness indexing needs b_global < w / 2^(n_scales-1)"); return -1;
  }
  launch_k(loss_finalize_kernel, dim3(1), dim3(1024), 0, as_stream(stream), *a);
  return finish_launch("loss_finalize");
}

}  // extern "C"
