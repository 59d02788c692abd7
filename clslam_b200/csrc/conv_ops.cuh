// Element-level definitions of the implicit-GEMM operands of the convolution stack.
// Everything here is __host__ __device__ so that tests/hostcheck can evaluate exactly the same
// index arithmetic on the CPU (test-only; the product path never runs these on the host).
#pragma once
#include "common.cuh"

namespace clslam {

// ---- forward / wgrad view of a layer's input: A[pixel][k], k = (r, s, ci), ci fastest ------
struct ConvIn {
  const float* x;    // PLAIN: (n, hi, wi, ci).  UPCAT: low-res part (n, h1, w1, c1)
  const float* x2;   // UPCAT: skip part (n, hi, wi, ci - c1), may be null when ci == c1
  int n, hi, wi, ci;
  int ho, wo;
  int kh, kw, stride, pad;
  int pad_mode, in_mode;
  int c1, h1, w1;
};

CLSLAM_HD ConvIn make_conv_in(const clslam_conv_desc& d, const float* x, const float* x2) {
  ConvIn c;
  c.x = x; c.x2 = x2;
  c.n = d.n; c.hi = d.hi; c.wi = d.wi; c.ci = d.ci; c.ho = d.ho; c.wo = d.wo;
  c.kh = d.kh; c.kw = d.kw; c.stride = d.stride; c.pad = d.pad;
  c.pad_mode = d.pad_mode; c.in_mode = d.in_mode; c.c1 = d.c1; c.h1 = d.h1; c.w1 = d.w1;
  return c;
}

// value of input channels [ch, ch+4) at input coordinate (ih, iw) AFTER border handling
CLSLAM_HD float4 conv_in_at(const ConvIn& c, int n, int ih, int iw, int ch) {
  if (c.in_mode == CLSLAM_IN_UPCAT) {
    if (ch < c.c1) {
      // legacy 'nearest': src = floor(dst * in / out)   (networks/depth_decoder.py:61,64)
      const int sh = (int)(((long long)ih * c.h1) / c.hi);
      const int sw = (int)(((long long)iw * c.w1) / c.wi);
      return ld4(c.x + (((size_t)n * c.h1 + sh) * c.w1 + sw) * c.c1 + ch);
    }
    const int c2 = c.ci - c.c1;
    return ld4(c.x2 + (((size_t)n * c.hi + ih) * c.wi + iw) * c2 + (ch - c.c1));
  }
  return ld4(c.x + (((size_t)n * c.hi + ih) * c.wi + iw) * c.ci + ch);
}

// A[pixel=(n,oh,ow)][k .. k+4)
CLSLAM_HD float4 conv_in_load(const ConvIn& c, int n, int oh, int ow, int k) {
  const int tap = k / c.ci;
  const int ch = k - tap * c.ci;
  const int r = tap / c.kw;
  const int s = tap - r * c.kw;
  int ih = oh * c.stride + r - c.pad;
  int iw = ow * c.stride + s - c.pad;
  if (c.pad_mode == CLSLAM_PAD_REFLECT) {
    ih = reflect_idx(ih, c.hi);
    iw = reflect_idx(iw, c.wi);
  } else if (ih < 0 || ih >= c.hi || iw < 0 || iw >= c.wi) {
    return f4zero();
  }
  return conv_in_at(c, n, ih, iw, ch);
}

// ---- split form of the same gather for the tensor-core stager: k = (r, s, ch) is decoded once per
// k block per thread (KTap), the pixel once per row (ih0 = oh*stride - pad, iw0 = ow*stride - pad).
struct KTap { int r, s, ch; };
CLSLAM_HD KTap ktap_decode(int k, int channels, int kw) {
  KTap t;
  const int tap = k / channels;
  t.ch = k - tap * channels;
  t.r = tap / kw;
  t.s = tap - t.r * kw;
  return t;
}
CLSLAM_HD float4 conv_in_gather(const ConvIn& c, int n, int ih0, int iw0, const KTap& t) {
  int ih = ih0 + t.r, iw = iw0 + t.s;
  if (c.pad_mode == CLSLAM_PAD_REFLECT) {
    ih = reflect_idx(ih, c.hi);
    iw = reflect_idx(iw, c.wi);
  } else if (ih < 0 || ih >= c.hi || iw < 0 || iw >= c.wi) {
    return f4zero();
  }
  if (c.in_mode == CLSLAM_IN_UPCAT) {
    if (t.ch < c.c1) {
      const int sh = (2 * c.h1 == c.hi) ? (ih >> 1) : (int)(((long long)ih * c.h1) / c.hi);
      const int sw = (2 * c.w1 == c.wi) ? (iw >> 1) : (int)(((long long)iw * c.w1) / c.wi);
      return ld4(c.x + (((size_t)n * c.h1 + sh) * c.w1 + sw) * c.c1 + t.ch);
    }
    return ld4(c.x2 + (((size_t)n * c.hi + ih) * c.wi + iw) * (c.ci - c.c1) + (t.ch - c.c1));
  }
  return ld4(c.x + (((size_t)n * c.hi + ih) * c.wi + iw) * c.ci + t.ch);
}

// ---- dgrad view: A[dX pixel][k], k = (r, s, co), co fastest ------------------------------
// dX[n,h,w,:] = sum over taps of W^T . dZ gathered at the output positions whose receptive
// field touches (h,w); with reflection padding the border rows/cols receive extra terms, and
// with UPCAT the low-res pixel receives the sum over its nearest-neighbour footprint.
struct DgradIn {
  const float* dz;   // (n, h, w, co) pre-activation gradient of this layer's output (stride 1)
  int n, h, w, co;
  int kh, kw, pad, pad_mode;
  int up;            // 1: dX is produced at (hl, wl) = footprint-summed
  int hl, wl;        // dX spatial dims (== h, w when !up and stride == 1)
  int stride;        // convolution stride (zero padding only when > 1): dZ is (h, w), dX is (hl, wl) ~ stride * (h, w)
};

// calls fn(src) for every dZ coordinate (1-D) that contributes through tap `r` to input
// coordinate `hh` (input and output have the same size: stride 1, pad = (k-1)/2)
template <class Fn>
CLSLAM_HD void dgrad_sources_1d(int hh, int r, int pad, int size, int reflect, Fn&& fn, int stride = 1) {
  const int base = hh + pad - r;
  if (stride > 1) {                    // output o reads input o * stride + r - pad: o = base / stride when divisible
    if (base >= 0 && base % stride == 0 && base / stride < size) fn(base / stride);
    return;
  }
  if (base >= 0 && base < size) fn(base);
  if (reflect) {                       // 3x3, pad 1 only
    if (hh == 1 && r == 0) fn(0);                    // padded row -1 mirrors row 1
    if (hh == size - 2 && r == 2) fn(size - 1);      // padded row `size` mirrors row size-2
  }
}

// footprint [lo, hi) of low-res index i under legacy nearest up-sampling low -> full
CLSLAM_HD void nearest_footprint(int i, int low, int full, int& lo, int& hi) {
  // src(d) = floor(d * low / full) == i  <=>  d in [ceil(i*full/low), ceil((i+1)*full/low))
  lo = (int)(((long long)i * full + low - 1) / low);
  hi = (int)(((long long)(i + 1) * full + low - 1) / low);
  if (hi > full) hi = full;
}

CLSLAM_HD float4 dgrad_in_load(const DgradIn& g, int n, int oh, int ow, int k) {
  const int tap = k / g.co;
  const int ch = k - tap * g.co;
  const int r = tap / g.kw;
  const int s = tap - r * g.kw;
  const int refl = g.pad_mode == CLSLAM_PAD_REFLECT;
  int h_lo = oh, h_hi = oh + 1, w_lo = ow, w_hi = ow + 1;
  if (g.up) {
    nearest_footprint(oh, g.hl, g.h, h_lo, h_hi);
    nearest_footprint(ow, g.wl, g.w, w_lo, w_hi);
  }
  float4 acc = f4zero();
  for (int hh = h_lo; hh < h_hi; ++hh) {
    dgrad_sources_1d(hh, r, g.pad, g.h, refl, [&](int sr) {
      for (int ww = w_lo; ww < w_hi; ++ww) {
        dgrad_sources_1d(ww, s, g.pad, g.w, refl, [&](int sc) {
          f4add(acc, ld4(g.dz + (((size_t)n * g.h + sr) * g.w + sc) * g.co + ch));
        }, g.stride);
      }
    }, g.stride);
  }
  return acc;
}

// split form for the tensor-core stager: footprint decoded once per row, tap once per k block
struct DgradRow { int n, h_lo, h_hi, w_lo, w_hi; };
CLSLAM_HD DgradRow dgrad_row(const DgradIn& g, int n, int oh, int ow) {
  DgradRow r{n, oh, oh + 1, ow, ow + 1};
  if (g.up) {
    nearest_footprint(oh, g.hl, g.h, r.h_lo, r.h_hi);
    nearest_footprint(ow, g.wl, g.w, r.w_lo, r.w_hi);
  }
  return r;
}
CLSLAM_HD float4 dgrad_in_gather(const DgradIn& g, const DgradRow& row, const KTap& t) {
  const int refl = g.pad_mode == CLSLAM_PAD_REFLECT;
  float4 acc = f4zero();
  for (int hh = row.h_lo; hh < row.h_hi; ++hh) {
    dgrad_sources_1d(hh, t.r, g.pad, g.h, refl, [&](int sr) {
      for (int ww = row.w_lo; ww < row.w_hi; ++ww) {
        dgrad_sources_1d(ww, t.s, g.pad, g.w, refl, [&](int sc) {
          f4add(acc, ld4(g.dz + (((size_t)row.n * g.h + sr) * g.w + sc) * g.co + t.ch));
        }, g.stride);
      }
    }, g.stride);
  }
  return acc;
}

// input gradient of a ONE-output-channel reflect 3x3 conv (dispconv) at pixel (n,h,w), channels
// [ch, ch+4): sum_taps W[r][s][ch..] * (sum of contributing dz).  dz: (n,1,h,w); w: [3][3][ci].
CLSLAM_HD float4 dispconv_dgrad_at(const float* dz, const float* wgt, int n, int hh, int ww, int hgt, int wid,
                                   int ci, int ch) {
  float4 acc = f4zero();
  for (int r = 0; r < 3; ++r) {
    for (int s = 0; s < 3; ++s) {
      float g = 0.f;
      dgrad_sources_1d(hh, r, 1, hgt, 1, [&](int sr) {
        dgrad_sources_1d(ww, s, 1, wid, 1, [&](int sc) { g += ld1(dz + ((size_t)n * hgt + sr) * wid + sc); });
      });
      const float4 wv = ld4(wgt + (r * 3 + s) * ci + ch);
      acc.x += g * wv.x; acc.y += g * wv.y; acc.z += g * wv.z; acc.w += g * wv.w;
    }
  }
  return acc;
}

}  // namespace clslam
