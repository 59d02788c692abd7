// Convolution stack of the adaptation hot path: forward, dgrad, wgrad as implicit GEMMs over the
// fp32 tile kernel in gemm_core.cuh.  See include/clslam_b200.h for the contract and the
// reference call sites each entry point replaces.
#include "conv_ops.cuh"
#include "gemm_core.cuh"
#include "gemm_tc.cuh"
#include "conv3x3_tc.cuh"
#include "planes.cuh"

namespace clslam {

// =========================== operand policies ===============================================

// A of forward: rows = output pixels, k = (r,s,ci)
struct AFwd {
  ConvIn c;
  int M, K;
  // element API used by the tcgen05 stager (gemm_tc.cuh)
  static constexpr bool kKContig = true;
  struct Row { int n, ih0, iw0; };
  using KState = KTap;
  __device__ Row decode(int m) const {
    Row r{-1, 0, 0};
    if (m >= 0 && m < M) {
      const int hw = c.ho * c.wo;
      r.n = m / hw;
      const int rem = m - r.n * hw;
      const int oh = rem / c.wo;
      r.ih0 = oh * c.stride - c.pad;
      r.iw0 = (rem - oh * c.wo) * c.stride - c.pad;
    }
    return r;
  }
  __device__ KState kdecode(int k) const { return ktap_decode(k, c.ci, c.kw); }
  __device__ float4 at(const Row& r, const KState& t) const { return r.n >= 0 ? conv_in_gather(c, r.n, r.ih0, r.iw0, t) : f4zero(); }
  template <int R> struct Frag {
    int n[KContigMap<R>::kVec], oh[KContigMap<R>::kVec], ow[KContigMap<R>::kVec];
    float4 r[KContigMap<R>::kVec];
  };
  template <int R> __device__ void init(Frag<R>& f, int m0, int tid) const {
    using Map = KContigMap<R>;
#pragma unroll
    for (int v = 0; v < Map::kVec; ++v) {
      const int m = m0 + Map::row(tid, v);
      if (Map::valid(tid, v) && m < M) {
        const int hw = c.ho * c.wo;
        f.n[v] = m / hw;
        const int rem = m - f.n[v] * hw;
        f.oh[v] = rem / c.wo;
        f.ow[v] = rem - f.oh[v] * c.wo;
      } else {
        f.n[v] = -1; f.oh[v] = 0; f.ow[v] = 0;
      }
    }
  }
  template <int R> __device__ void fetch(Frag<R>& f, int k0, int k_end, int tid) const {
    using Map = KContigMap<R>;
#pragma unroll
    for (int v = 0; v < Map::kVec; ++v) {
      const int k = k0 + Map::k4(tid, v) * 4;
      f.r[v] = (f.n[v] >= 0 && k < k_end) ? conv_in_load(c, f.n[v], f.oh[v], f.ow[v], k) : f4zero();
    }
  }
  template <int R> __device__ void commit(const Frag<R>& f, float* S, int ld, int tid) const {
    KContigMap<R>::commit(S, ld, tid, f.r);
  }
};

// B of forward: rows = output channels, W[co][K] K-contiguous
struct BWeightK {
  const float* w;
  int N, K;
  static constexpr bool kKContig = true;
  struct Row { const float* p; };
  using KState = int;
  __device__ Row decode(int n) const { return Row{(n >= 0 && n < N) ? w + (size_t)n * K : nullptr}; }
  __device__ KState kdecode(int k) const { return k; }
  __device__ float4 at(const Row& r, const KState& k) const { return r.p ? ld4(r.p + k) : f4zero(); }
  template <int R> struct Frag { const float* p[KContigMap<R>::kVec]; float4 r[KContigMap<R>::kVec]; };
  template <int R> __device__ void init(Frag<R>& f, int n0, int tid) const {
    using Map = KContigMap<R>;
#pragma unroll
    for (int v = 0; v < Map::kVec; ++v) {
      const int n = n0 + Map::row(tid, v);
      f.p[v] = (Map::valid(tid, v) && n < N) ? w + (size_t)n * K : nullptr;
    }
  }
  template <int R> __device__ void fetch(Frag<R>& f, int k0, int k_end, int tid) const {
    using Map = KContigMap<R>;
#pragma unroll
    for (int v = 0; v < Map::kVec; ++v) {
      const int k = k0 + Map::k4(tid, v) * 4;
      f.r[v] = (f.p[v] != nullptr && k < k_end) ? ld4(f.p[v] + k) : f4zero();
    }
  }
  template <int R> __device__ void commit(const Frag<R>& f, float* S, int ld, int tid) const {
    KContigMap<R>::commit(S, ld, tid, f.r);
  }
};

// A of dgrad: rows = dX pixels, k = (r,s,co)
struct ADgrad {
  DgradIn g;
  int M, K;
  static constexpr bool kKContig = true;
  using Row = DgradRow;
  using KState = KTap;
  __device__ Row decode(int m) const {
    if (m < 0 || m >= M) return Row{-1, 0, 0, 0, 0};
    const int hw = g.hl * g.wl;
    const int n = m / hw;
    const int rem = m - n * hw;
    const int oh = rem / g.wl;
    return dgrad_row(g, n, oh, rem - oh * g.wl);
  }
  __device__ KState kdecode(int k) const { return ktap_decode(k, g.co, g.kw); }
  __device__ float4 at(const Row& r, const KState& t) const { return r.n >= 0 ? dgrad_in_gather(g, r, t) : f4zero(); }
  template <int R> struct Frag {
    int n[KContigMap<R>::kVec], oh[KContigMap<R>::kVec], ow[KContigMap<R>::kVec];
    float4 r[KContigMap<R>::kVec];
  };
  template <int R> __device__ void init(Frag<R>& f, int m0, int tid) const {
    using Map = KContigMap<R>;
#pragma unroll
    for (int v = 0; v < Map::kVec; ++v) {
      const int m = m0 + Map::row(tid, v);
      if (Map::valid(tid, v) && m < M) {
        const int hw = g.hl * g.wl;
        f.n[v] = m / hw;
        const int rem = m - f.n[v] * hw;
        f.oh[v] = rem / g.wl;
        f.ow[v] = rem - f.oh[v] * g.wl;
      } else {
        f.n[v] = -1; f.oh[v] = 0; f.ow[v] = 0;
      }
    }
  }
  template <int R> __device__ void fetch(Frag<R>& f, int k0, int k_end, int tid) const {
    using Map = KContigMap<R>;
#pragma unroll
    for (int v = 0; v < Map::kVec; ++v) {
      const int k = k0 + Map::k4(tid, v) * 4;
      f.r[v] = (f.n[v] >= 0 && k < k_end) ? dgrad_in_load(g, f.n[v], f.oh[v], f.ow[v], k) : f4zero();
    }
  }
  template <int R> __device__ void commit(const Frag<R>& f, float* S, int ld, int tid) const {
    KContigMap<R>::commit(S, ld, tid, f.r);
  }
};

// B of dgrad: B[k=(tap,co)][n=ci] = W[co][tap][ci]  (ci contiguous)
struct BWeightDgrad {
  const float* w;
  int N;        // number of input channels produced (c1 for UPCAT)
  int ci_full;  // row length of the weight tensor
  int co, taps;
};

}  // namespace clslam

// The B-of-dgrad and the wgrad operands need the tile origin at fetch time; to keep one kernel
// signature they store it in the Frag at init().
namespace clslam {

struct BWeightDgradOp {
  BWeightDgrad b;
  static constexpr bool kKContig = false;
  using RowState = int;                 // input-channel index, -1 = out of range
  using KState = const float*;          // row of W for this (tap, co)
  __device__ RowState rdecode(int n) const { return n < b.N ? n : -1; }
  __device__ KState kdecode(int k) const {
    const int tap = k / b.co, o = k - tap * b.co;
    return b.w + ((size_t)o * b.taps + tap) * b.ci_full;
  }
  __device__ float4 at4(const KState& p, const RowState& n) const { return n >= 0 ? ld4(p + n) : f4zero(); }
  template <int R> struct Frag { int n0; float4 r[RContigMap<R>::kVec]; };
  template <int R> __device__ void init(Frag<R>& f, int n0, int) const { f.n0 = n0; }
  template <int R> __device__ void fetch(Frag<R>& f, int k0, int k_end, int tid) const {
    using Map = RContigMap<R>;
#pragma unroll
    for (int v = 0; v < Map::kVec; ++v) {
      const int k = k0 + Map::kk(tid, v);
      const int n = f.n0 + Map::r4(tid, v);
      if (Map::valid(tid, v) && k < k_end && n < b.N) {
        const int tap = k / b.co, o = k - tap * b.co;
        f.r[v] = ld4(b.w + ((size_t)o * b.taps + tap) * b.ci_full + n);
      } else {
        f.r[v] = f4zero();
      }
    }
  }
  template <int R> __device__ void commit(const Frag<R>& f, float* S, int ld, int tid) const {
    RContigMap<R>::commit(S, ld, tid, f.r);
  }
};

// A of wgrad: A[m=co][k=pixel] = dz[pixel][co]   (co contiguous)
struct AWgradDz {
  const float* dz;
  int co, P;
  static constexpr bool kKContig = false;
  using RowState = int;
  using KState = const float*;
  __device__ RowState rdecode(int m) const { return m < co ? m : -1; }
  __device__ KState kdecode(int p) const { return dz + (size_t)p * co; }
  __device__ float4 at4(const KState& p, const RowState& m) const { return m >= 0 ? ld4(p + m) : f4zero(); }
  template <int R> struct Frag { int m0; float4 r[RContigMap<R>::kVec]; };
  template <int R> __device__ void init(Frag<R>& f, int m0, int) const { f.m0 = m0; }
  template <int R> __device__ void fetch(Frag<R>& f, int k0, int k_end, int tid) const {
    using Map = RContigMap<R>;
#pragma unroll
    for (int v = 0; v < Map::kVec; ++v) {
      const int p = k0 + Map::kk(tid, v);
      const int m = f.m0 + Map::r4(tid, v);
      f.r[v] = (Map::valid(tid, v) && p < k_end && m < co) ? ld4(dz + (size_t)p * co + m) : f4zero();
    }
  }
  template <int R> __device__ void commit(const Frag<R>& f, float* S, int ld, int tid) const {
    RContigMap<R>::commit(S, ld, tid, f.r);
  }
};

// B of wgrad: B[k=pixel][n=(r,s,ci)] = gathered layer input
struct BWgradIn {
  ConvIn c;
  int Kw;   // kh*kw*ci
  static constexpr bool kKContig = false;
  using RowState = KTap;                // (r, s, ch) of this weight column; r = -1: out of range
  struct KState { int n, ih0, iw0; };   // the pixel
  // column Kw is a virtual all-ones column: its GEMM output is the bias gradient sum_p dz[p][co]
  __device__ RowState rdecode(int kidx) const {
    return kidx < Kw ? ktap_decode(kidx, c.ci, c.kw) : (kidx == Kw ? KTap{-2, 0, 0} : KTap{-1, 0, 0});
  }
  __device__ KState kdecode(int p) const {
    const int hw = c.ho * c.wo;
    const int n = p / hw;
    const int rem = p - n * hw;
    const int oh = rem / c.wo;
    return KState{n, oh * c.stride - c.pad, (rem - oh * c.wo) * c.stride - c.pad};
  }
  __device__ float4 at4(const KState& px, const RowState& t) const {
    if (t.r >= 0) return conv_in_gather(c, px.n, px.ih0, px.iw0, t);
    return t.r == -2 ? make_float4(1.f, 0.f, 0.f, 0.f) : f4zero();
  }
  template <int R> struct Frag { int n0; float4 r[RContigMap<R>::kVec]; };
  template <int R> __device__ void init(Frag<R>& f, int n0, int) const { f.n0 = n0; }
  template <int R> __device__ void fetch(Frag<R>& f, int k0, int k_end, int tid) const {
    using Map = RContigMap<R>;
#pragma unroll
    for (int v = 0; v < Map::kVec; ++v) {
      const int p = k0 + Map::kk(tid, v);
      const int kidx = f.n0 + Map::r4(tid, v);
      if (Map::valid(tid, v) && p < k_end && kidx < Kw) {
        const int hw = c.ho * c.wo;
        const int n = p / hw;
        const int rem = p - n * hw;
        const int oh = rem / c.wo;
        f.r[v] = conv_in_load(c, n, oh, rem - oh * c.wo, kidx);
      } else {
        f.r[v] = f4zero();
      }
    }
  }
  template <int R> __device__ void commit(const Frag<R>& f, float* S, int ld, int tid) const {
    RContigMap<R>::commit(S, ld, tid, f.r);
  }
};

// =========================== epilogues ======================================================

struct EpiFwd {
  static constexpr bool kBias = false;
  float* y; const float* scale; const float* shift; const float* res;
  int N, act;
  PlaneView po;       // optional second output: bf16 hi/lo planes (base == nullptr: none)
  int ho, wo;         // output spatial dims (row m -> (image, oh, ow)) when po is set
  __device__ void emit(int m, int n, const float (&o)[4]) const {
    const int hw = ho * wo;
    const int img = m / hw, rem = m - img * hw;
    const int oh = rem / wo;
    planes_emit<4>(po, img, oh, rem - oh * wo, n, o);
  }
  // tensor-core kernel: per-column parameters are staged in shared memory once per CTA
  __device__ void tc_prologue(float* sp, int n0, int bn, int n_total, int tid) const {
    for (int j = tid; j < bn; j += kTcThreads) {
      const int n = n0 + j;
      sp[j] = (scale && n < n_total) ? scale[n] : 1.f;
      sp[bn + j] = (shift && n < n_total) ? shift[n] : 0.f;
    }
  }
  __device__ void tc_store(int m, int n, const float (&v)[4], const float* sp, int nl, int bn) const {
    const size_t off = (size_t)m * N + n;
    float4 r4 = f4zero();
    if (res) r4 = *reinterpret_cast<const float4*>(res + off);
    const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = act_apply(v[j] * sp[nl + j] + sp[bn + nl + j] + rr[j], act);
    if (y) *reinterpret_cast<float4*>(y + off) = make_float4(o[0], o[1], o[2], o[3]);
    if (po.base) emit(m, n, o);
  }
  __device__ void store(int m, int n, const float (&v)[4]) const {
    float o[4];
    const size_t off = (size_t)m * N + n;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = v[j];
      if (scale) t *= scale[n + j];
      if (shift) t += shift[n + j];
      if (res) t += res[off + j];
      o[j] = act_apply(t, act);
    }
    if (y) *reinterpret_cast<float4*>(y + off) = make_float4(o[0], o[1], o[2], o[3]);
    if (po.base) emit(m, n, o);
  }
  __device__ void store_bias(int, float) const {}
};

struct EpiDgrad {
  static constexpr bool kBias = false;
  __device__ void tc_prologue(float*, int, int, int, int) const {}
  __device__ void tc_store(int m, int n, const float (&v)[4], const float*, int, int) const { store(m, n, v); }
  float* dzp; const float* y_prev; int act_prev;
  const float* disp_dz; const float* disp_w;
  int N;          // channels of dX
  int h, w;       // dX spatial dims (for the dispconv term)
  __device__ void store(int m, int n, const float (&v)[4]) const {
    float o[4] = {v[0], v[1], v[2], v[3]};
    if (disp_dz) {
      const int hw = h * w;
      const int b = m / hw, rem = m - b * hw;
      const int hh = rem / w, ww = rem - hh * w;
      const float4 d = dispconv_dgrad_at(disp_dz, disp_w, b, hh, ww, h, w, N, n);
      o[0] += d.x; o[1] += d.y; o[2] += d.z; o[3] += d.w;
    }
    const size_t off = (size_t)m * N + n;
    if (y_prev) {
      const float4 yv = *reinterpret_cast<const float4*>(y_prev + off);
      o[0] *= act_grad_from_output(yv.x, act_prev);
      o[1] *= act_grad_from_output(yv.y, act_prev);
      o[2] *= act_grad_from_output(yv.z, act_prev);
      o[3] *= act_grad_from_output(yv.w, act_prev);
    }
    *reinterpret_cast<float4*>(dzp + off) = make_float4(o[0], o[1], o[2], o[3]);
  }
  __device__ void store_bias(int, float) const {}
};

struct EpiWgrad {
  static constexpr bool kBias = true;
  __device__ void tc_prologue(float*, int, int, int, int) const {}
  __device__ void tc_store(int m, int n, const float (&v)[4], const float*, int, int) const { store(m, n, v); }
  float* dw; float* db; int N;
  __device__ void store(int m, int n, const float (&v)[4]) const {
    *reinterpret_cast<float4*>(dw + (size_t)m * N + n) = make_float4(v[0], v[1], v[2], v[3]);
  }
  __device__ void store_bias(int m, float v) const { if (db) db[m] = v; }
};

// tensor-core wgrad: N' = Kw + 4, column Kw carries the bias gradient (see BWgradIn::rdecode)
struct EpiWgradTc {
  static constexpr bool kBias = false;
  __device__ void tc_prologue(float*, int, int, int, int) const {}
  __device__ void tc_store(int m, int n, const float (&v)[4], const float*, int, int) const { store(m, n, v); }
  float* dw; float* db; int Kw;
  __device__ void store(int m, int n, const float (&v)[4]) const {
    if (n < Kw) *reinterpret_cast<float4*>(dw + (size_t)m * Kw + n) = make_float4(v[0], v[1], v[2], v[3]);
    else if (n == Kw && db) db[m] = v[0];
  }
  __device__ void store_bias(int, float) const {}
};

// =========================== launch plumbing ================================================

struct TilePlan { int cfg; int bm, bn; int gx, gy, splits, k_chunk; };

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_num_sms <= 0)
      g_num_sms = 148;
  }
  return g_num_sms;
}

// cfg: 0 = 128x64, 1 = 64x64, 2 = 128x32, 3 = 256x16, 4 = 32x64, 5 = 16x64
static TilePlan plan_tiles(int M, int N, int K, bool allow_split) {
  TilePlan p{};
  if (N <= 16) { p.cfg = (M <= 16) ? 5 : 3; }
  else if (N <= 32) { p.cfg = (M <= 32) ? 4 : 2; }
  else if (M <= 16) p.cfg = 5;
  else if (M <= 32) p.cfg = 4;
  else if (M <= 64 || (long long)((M + 127) / 128) * ((N + 63) / 64) < num_sms()) p.cfg = (M <= 32) ? 4 : 1;
  else p.cfg = 0;
  static const int bms[6] = {128, 64, 128, 256, 32, 16};
  static const int bns[6] = {64, 64, 32, 16, 64, 64};
  p.bm = bms[p.cfg]; p.bn = bns[p.cfg];
  if (p.cfg == 1 && (long long)((M + 63) / 64) * ((N + 63) / 64) * 4 < num_sms() && M <= 256) { p.cfg = 4; p.bm = 32; }
  p.gx = (M + p.bm - 1) / p.bm;
  p.gy = (N + p.bn - 1) / p.bn;
  p.splits = 1;
  const int k_tiles = (K + kBK - 1) / kBK;
  if (allow_split) {
    const long long tiles = (long long)p.gx * p.gy;
    const int target = 2 * num_sms();
    if (tiles < target && k_tiles >= 8) {
      int s = (int)((target + tiles - 1) / tiles);
      const int max_s = k_tiles / 4;            // at least 4 k-tiles (64 k) per split
      if (s > max_s) s = max_s;
      if (s > 64) s = 64;
      if (s < 1) s = 1;
      p.splits = s;
    }
  }
  const int tiles_per_split = (k_tiles + p.splits - 1) / p.splits;
  p.k_chunk = tiles_per_split * kBK;
  p.splits = (k_tiles + tiles_per_split - 1) / tiles_per_split;
  return p;
}

template <class AOp, class BOp, class Epi>
static int launch_gemm(const TilePlan& p, const AOp& a, const BOp& b, const Epi& e, int M, int N, int K,
                       float* ws, long long ws_bytes, cudaStream_t st, const char* what) {
  const bool has_bias = Epi::kBias;
  if (p.splits > 1) {
    const long long need = ((long long)p.splits * M * N + (has_bias ? (long long)p.splits * M : 0)) * 4;
    if (ws == nullptr || ws_bytes < need) { set_error("%s: workspace too small (%lld < %lld)", what, ws_bytes, need); return -2; }
  }
  dim3 grid(p.gx, p.gy, p.splits), block(kGemmThreads);
  switch (p.cfg) {
    case 0: launch_k(gemm_tile_kernel<128, 64, 8, AOp, BOp, Epi>, dim3(grid), dim3(block), 0, st, a, b, e, M, N, K, p.k_chunk, ws); break;
    case 1: launch_k(gemm_tile_kernel<64, 64, 4, AOp, BOp, Epi>, dim3(grid), dim3(block), 0, st, a, b, e, M, N, K, p.k_chunk, ws); break;
    case 2: launch_k(gemm_tile_kernel<128, 32, 4, AOp, BOp, Epi>, dim3(grid), dim3(block), 0, st, a, b, e, M, N, K, p.k_chunk, ws); break;
    case 3: launch_k(gemm_tile_kernel<256, 16, 4, AOp, BOp, Epi>, dim3(grid), dim3(block), 0, st, a, b, e, M, N, K, p.k_chunk, ws); break;
    case 4: launch_k(gemm_tile_kernel<32, 64, 2, AOp, BOp, Epi>, dim3(grid), dim3(block), 0, st, a, b, e, M, N, K, p.k_chunk, ws); break;
    default: launch_k(gemm_tile_kernel<16, 64, 1, AOp, BOp, Epi>, dim3(grid), dim3(block), 0, st, a, b, e, M, N, K, p.k_chunk, ws); break;
  }
  int rc = finish_launch(what);
  if (rc) return rc;
  if (p.splits > 1) {
    const long long total4 = (long long)M * N / 4;
    int blocks = (int)((total4 + 255) / 256);
    if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
    if (blocks < 1) blocks = 1;
    launch_k(splitk_reduce_kernel<Epi>, dim3(blocks), dim3(256), 0, st, e, ws, M, N, p.splits);
    rc = finish_launch(what);
  }
  return rc;
}

// ---- tensor-core path -----------------------------------------------------------------------
static int g_conv_mode = 1;      // 0 = exact fp32 CUDA-core path, 1 = tcgen05 3xBF16

struct TcPlan { int bn; int gx, gy, splits, k_chunk; };

static TcPlan plan_tc(int M, int N, int K) {
  TcPlan p{};
  p.bn = N >= 128 ? 128 : (N > 32 ? 64 : (N > 16 ? 32 : 16));
  p.gx = (M + kTcBM - 1) / kTcBM;
  p.gy = (N + p.bn - 1) / p.bn;
  const int k_tiles = (K + kTcBK - 1) / kTcBK;
  const long long tiles = (long long)p.gx * p.gy;
  int s = 1;
  const int target = num_sms();
  if (tiles < target && k_tiles >= 4) {
    s = (int)((target + tiles - 1) / tiles);
    const int max_s = k_tiles / 2;              // at least 2 k-blocks (128 k) per split
    if (s > max_s) s = max_s;
    const int cap = ((long long)M * N <= 65536) ? 128 : 32;     // small outputs (late-decoder wgrad): partials are cheap
    if (s > cap) s = cap;
    if (s < 1) s = 1;
  }
  const int per = (k_tiles + s - 1) / s;
  p.k_chunk = per * kTcBK;
  p.splits = (k_tiles + per - 1) / per;
  return p;
}

template <int BN, class AOp, class BOp, class Epi>
static int launch_tc_bn(const TcPlan& p, const AOp& a, const BOp& b, const Epi& e, int M, int N, int K, float* ws,
                        cudaStream_t st, const char* what) {
  auto kern = gemm_tc_kernel<BN, AOp, BOp, Epi>;
  constexpr int kSmem = TcSmem<BN, AOp, BOp>::kTotal;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (err != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(err)); return -3; }
    attr_set = true;
  }
  dim3 grid(p.gx, p.gy, p.splits);
  launch_k(kern, dim3(grid), dim3(kTcBlock), kSmem, st, a, b, e, M, N, K, p.k_chunk, ws);
  return finish_launch(what);
}

template <class AOp, class BOp, class Epi>
static int launch_gemm_tc(const TcPlan& p, const AOp& a, const BOp& b, const Epi& e, int M, int N, int K, float* ws,
                          long long ws_bytes, cudaStream_t st, const char* what) {
  if (p.splits > 1) {
    const long long need = ((long long)p.splits * M * N + (Epi::kBias ? (long long)p.splits * M : 0)) * 4;
    if (ws == nullptr || ws_bytes < need) { set_error("%s: workspace too small (%lld < %lld)", what, ws_bytes, need); return -2; }
  }
  int rc;
  switch (p.bn) {
    case 128: rc = launch_tc_bn<128>(p, a, b, e, M, N, K, ws, st, what); break;
    case 64: rc = launch_tc_bn<64>(p, a, b, e, M, N, K, ws, st, what); break;
    case 32: rc = launch_tc_bn<32>(p, a, b, e, M, N, K, ws, st, what); break;
    default: rc = launch_tc_bn<16>(p, a, b, e, M, N, K, ws, st, what); break;
  }
  if (rc) return rc;
  if (p.splits > 1) {
    const long long total4 = (long long)M * N / 4;
    int blocks = (int)((total4 + 255) / 256);
    if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
    if (blocks < 1) blocks = 1;
    launch_k(splitk_reduce_kernel<Epi>, dim3(blocks), dim3(256), 0, st, e, ws, M, N, p.splits);
    rc = finish_launch(what);
  }
  return rc;
}

// workspace need of either path (the larger, so that the mode can be switched at run time)
static long long ws_need(int M, int N, int K, bool bias) {
  const TilePlan p0 = plan_tiles(M, N, K, true);
  const int n_tc = bias ? N + 4 : N;                       // tensor-core wgrad carries the bias in 4 extra columns
  const TcPlan p1 = plan_tc(M, n_tc, K);
  const long long a = p0.splits > 1 ? ((long long)p0.splits * M * N + (bias ? (long long)p0.splits * M : 0)) * 4 : 0;
  const long long b = p1.splits > 1 ? (long long)p1.splits * M * n_tc * 4 : 0;
  return a > b ? a : b;
}

template <class AOp, class BOp, class Epi>
static int run_gemm(const AOp& a, const BOp& b, const Epi& e, int M, int N, int K, float* ws, long long ws_bytes,
                    cudaStream_t st, const char* what) {
  if (g_conv_mode == 1 && K >= 32)
    return launch_gemm_tc(plan_tc(M, N, K), a, b, e, M, N, K, ws, ws_bytes, st, what);
  return launch_gemm(plan_tiles(M, N, K, true), a, b, e, M, N, K, ws, ws_bytes, st, what);
}

// ---- 3x3 patch kernel (conv3x3_tc.cuh) -------------------------------------------------------
struct P3Plan { int bn, gx, gy, splits; P3Geom g; };

static bool p3_eligible(const clslam_conv_desc* d) {
  if (d->kh != 3 || d->kw != 3 || d->stride != 1 || d->pad != 1 || d->ci % 64 != 0) return false;
  if (d->hi != d->ho || d->wi != d->wo) return false;
  if (d->in_mode == CLSLAM_IN_UPCAT && d->c1 % 64 != 0) return false;
  // the 16 x 8 pixel M tiles must be reasonably full (6 x 20 maps waste 2/3 of every tile: generic kernel instead)
  const long long tiles = (long long)((d->wo + kP3TW - 1) / kP3TW) * ((d->ho + kP3TH - 1) / kP3TH);
  return (long long)d->ho * d->wo * 10 >= tiles * kTcBM * 6;
}

static P3Plan plan_p3(const clslam_conv_desc* d) {
  P3Plan p{};
  p.bn = d->co >= 128 ? 128 : (d->co > 32 ? 64 : (d->co > 16 ? 32 : 16));
  p.g.tiles_x = (d->wo + kP3TW - 1) / kP3TW;
  p.g.tiles_y = (d->ho + kP3TH - 1) / kP3TH;
  p.g.n_chunks = d->ci / 64;
  p.gx = d->n * p.g.tiles_x * p.g.tiles_y;
  p.gy = (d->co + p.bn - 1) / p.bn;
  int s = 1;
  const long long tiles = (long long)p.gx * p.gy;
  if (tiles < num_sms() && p.g.n_chunks > 1) {
    s = (int)((num_sms() + tiles - 1) / tiles);
    if (s > p.g.n_chunks) s = p.g.n_chunks;
  }
  p.g.chunks_per_split = (p.g.n_chunks + s - 1) / s;
  p.splits = (p.g.n_chunks + p.g.chunks_per_split - 1) / p.g.chunks_per_split;
  return p;
}

template <int BN>
static int launch_p3(const P3Plan& p, const ConvIn& c, const uint16_t* w_hi, const uint16_t* w_lo, const EpiFwd& e, int co,
                     float* ws, cudaStream_t st) {
  auto kern = conv3x3_tc_kernel<BN, EpiFwd>;
  constexpr int kSmem = P3Smem<BN>::kTotal;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (err != cudaSuccess) { set_error("conv3x3_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(err)); return -3; }
    attr_set = true;
  }
  dim3 grid(p.gx, p.gy, p.splits);
  launch_k(kern, dim3(grid), dim3(kTcBlock), kSmem, st, c, w_hi, w_lo, e, co, p.g, ws);
  return finish_launch("conv3x3_tc");
}

static int check_desc(const clslam_conv_desc* d, const char* what) {
  if (!d) { set_error("%s: null descriptor", what); return -1; }
  if (d->ci % 4 != 0 || d->co % 4 != 0) { set_error("%s: channels must be multiples of 4 (ci=%d co=%d)", what, d->ci, d->co); return -1; }
  if (d->in_mode == CLSLAM_IN_UPCAT && (d->c1 % 4 != 0 || d->c1 <= 0 || d->c1 > d->ci || d->stride != 1)) { set_error("%s: bad UPCAT", what); return -1; }
  if (d->pad_mode == CLSLAM_PAD_REFLECT && !(d->kh == 3 && d->kw == 3 && d->pad == 1 && d->stride == 1 && d->hi >= 2 && d->wi >= 2)) { set_error("%s: reflect padding needs 3x3/s1/p1", what); return -1; }
  const int eh = (d->hi + 2 * d->pad - d->kh) / d->stride + 1, ew = (d->wi + 2 * d->pad - d->kw) / d->stride + 1;
  if (eh != d->ho || ew != d->wo) { set_error("%s: output size mismatch (%dx%d vs %dx%d)", what, d->ho, d->wo, eh, ew); return -1; }
  return 0;
}

}  // namespace clslam

using namespace clslam;

extern "C" {

int clslam_set_conv_mode(int mode) {
  if (mode != 0 && mode != 1) { set_error("set_conv_mode: 0 (fp32 CUDA cores) or 1 (tcgen05 3xBF16)"); return -1; }
  g_conv_mode = mode;
  return 0;
}
int clslam_get_conv_mode(void) { return g_conv_mode; }
int clslam_debug_tc_stamps(unsigned long long* out16) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out16, g_tc_stamps, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -1;
}

long long clslam_conv_fwd_workspace_bytes(const clslam_conv_desc* d) {
  if (check_desc(d, "conv_fwd_ws")) return -1;
  const int M = d->n * d->ho * d->wo, N = d->co, K = d->kh * d->kw * d->ci;
  long long need = ws_need(M, N, K, false);
  if (p3_eligible(d)) {
    const P3Plan p = plan_p3(d);
    const long long n3 = p.splits > 1 ? (long long)p.splits * M * N * 4 : 0;
    if (n3 > need) need = n3;
  }
  return need;
}

int clslam_split_bf16(const float* src, void* hi, void* lo, long long n, void* stream) {
  if (n % 4) { set_error("split_bf16: n %% 4 != 0"); return -1; }
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  launch_k(split_bf16_kernel, dim3((int)blocks), dim3(256), 0, as_stream(stream), src, (uint16_t*)hi, (uint16_t*)lo, n / 4);
  return finish_launch("split_bf16");
}

static int make_po(const clslam_conv_desc* d, const clslam_planes* out, int up, int halo, PlaneView& po, const char* what) {
  po = PlaneView{};
  if (!out) return 0;
  if (!out->base || (up != 1 && up != 2) || out->h != d->ho * up || out->w != d->wo * up || out->c != d->co || out->n != d->n ||
      out->c_total % 8 || out->c_off % 4 || out->c_off + out->c > out->c_total) {
    set_error("%s: plane output does not match the layer (%d x %d x %d, up %d)", what, d->ho, d->wo, d->co, up);
    return -1;
  }
  po = make_plane_view(out, up, halo);
  return 0;
}

static int conv_fwd_generic(const clslam_conv_desc* d, const float* x, const float* x2, const float* w, const float* scale,
                            const float* shift, const float* residual, float* y, const PlaneView& po, void* workspace,
                            long long workspace_bytes, void* stream) {
  const int M = d->n * d->ho * d->wo, N = d->co, K = d->kh * d->kw * d->ci;
  AFwd a{make_conv_in(*d, x, x2), M, K};
  BWeightK b{w, N, K};
  EpiFwd e{y, scale, shift, residual, N, d->act, po, d->ho, d->wo};
  return run_gemm(a, b, e, M, N, K, (float*)workspace, workspace_bytes, as_stream(stream), "conv_fwd");
}

int clslam_conv_fwd_emit(const clslam_conv_desc* d, const float* x, const float* x2, const float* w, const void* w_hi,
                         const void* w_lo, const float* scale, const float* shift, const float* residual, float* y,
                         const clslam_planes* out, int out_up, int out_halo, void* workspace, long long workspace_bytes,
                         void* stream) {
  if (int rc = check_desc(d, "conv_fwd")) return rc;
  PlaneView po;
  if (int rc = make_po(d, out, out_up, out_halo, po, "conv_fwd")) return rc;
  if (!y && !po.base) { set_error("conv_fwd: no output"); return -1; }
  if (g_conv_mode != 1 || !w_hi || !w_lo || !p3_eligible(d))
    return conv_fwd_generic(d, x, x2, w, scale, shift, residual, y, po, workspace, workspace_bytes, stream);
  const int M = d->n * d->ho * d->wo, N = d->co;
  const P3Plan p = plan_p3(d);
  if (p.splits > 1 && (workspace == nullptr || workspace_bytes < (long long)p.splits * M * N * 4)) {
    set_error("conv_fwd_bf16w: workspace too small"); return -2;
  }
  EpiFwd e{y, scale, shift, residual, N, d->act, po, d->ho, d->wo};
  const ConvIn c = make_conv_in(*d, x, x2);
  cudaStream_t st = as_stream(stream);
  int rc;
  switch (p.bn) {
    case 128: rc = launch_p3<128>(p, c, (const uint16_t*)w_hi, (const uint16_t*)w_lo, e, N, (float*)workspace, st); break;
    case 64: rc = launch_p3<64>(p, c, (const uint16_t*)w_hi, (const uint16_t*)w_lo, e, N, (float*)workspace, st); break;
    case 32: rc = launch_p3<32>(p, c, (const uint16_t*)w_hi, (const uint16_t*)w_lo, e, N, (float*)workspace, st); break;
    default: rc = launch_p3<16>(p, c, (const uint16_t*)w_hi, (const uint16_t*)w_lo, e, N, (float*)workspace, st); break;
  }
  if (rc) return rc;
  if (p.splits > 1) {
    const long long total4 = (long long)M * N / 4;
    int blocks = (int)((total4 + 255) / 256);
    if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
    if (blocks < 1) blocks = 1;
    launch_k(splitk_reduce_kernel<EpiFwd>, dim3(blocks), dim3(256), 0, st, e, (const float*)workspace, M, N, p.splits);
    rc = finish_launch("conv_fwd_bf16w");
  }
  return rc;
}

int clslam_conv_fwd_bf16w(const clslam_conv_desc* d, const float* x, const float* x2, const float* w, const void* w_hi,
                          const void* w_lo, const float* scale, const float* shift, const float* residual, float* y,
                          void* workspace, long long workspace_bytes, void* stream) {
  return clslam_conv_fwd_emit(d, x, x2, w, w_hi, w_lo, scale, shift, residual, y, nullptr, 1, 0, workspace, workspace_bytes, stream);
}

int clslam_conv_fwd(const clslam_conv_desc* d, const float* x, const float* x2, const float* w, const float* scale,
                    const float* shift, const float* residual, float* y, void* workspace, long long workspace_bytes,
                    void* stream) {
  if (int rc = check_desc(d, "conv_fwd")) return rc;
  return conv_fwd_generic(d, x, x2, w, scale, shift, residual, y, PlaneView{}, workspace, workspace_bytes, stream);
}

static void dgrad_dims(const clslam_conv_desc* d, int& M, int& N, int& K, DgradIn& g, const float* dz) {
  g.dz = dz; g.n = d->n; g.h = d->ho; g.w = d->wo; g.co = d->co;
  g.kh = d->kh; g.kw = d->kw; g.pad = d->pad; g.pad_mode = d->pad_mode;
  g.up = d->in_mode == CLSLAM_IN_UPCAT;
  g.stride = d->stride;
  g.hl = g.up ? d->h1 : d->hi; g.wl = g.up ? d->w1 : d->wi;
  M = d->n * g.hl * g.wl;
  N = g.up ? d->c1 : d->ci;
  K = d->kh * d->kw * d->co;
}

long long clslam_conv_dgrad_workspace_bytes(const clslam_conv_desc* d) {
  if (check_desc(d, "conv_dgrad_ws")) return -1;
  int M, N, K; DgradIn g; dgrad_dims(d, M, N, K, g, nullptr);
  return ws_need(M, N, K, false);
}

int clslam_conv_dgrad(const clslam_conv_desc* d, const float* dz, const float* w, const float* y_prev, int act_prev,
                      const float* disp_dz, const float* disp_w, float* dz_prev, void* workspace,
                      long long workspace_bytes, void* stream) {
  if (int rc = check_desc(d, "conv_dgrad")) return rc;
  if (d->stride != 1 || d->hi != d->ho || d->wi != d->wo) { set_error("conv_dgrad: only stride-1 'same' convolutions"); return -1; }
  if (disp_dz && d->in_mode == CLSLAM_IN_UPCAT) { set_error("conv_dgrad: dispconv term needs a PLAIN input"); return -1; }
  int M, N, K; DgradIn g; dgrad_dims(d, M, N, K, g, dz);
  ADgrad a{g, M, K};
  BWeightDgradOp b{{w, N, d->ci, d->co, d->kh * d->kw}};
  EpiDgrad e{dz_prev, y_prev, act_prev, disp_dz, disp_w, N, g.hl, g.wl};
  return run_gemm(a, b, e, M, N, K, (float*)workspace, workspace_bytes, as_stream(stream), "conv_dgrad");
}

int clslam_conv_dgrad_ex(const clslam_conv_desc* d, const float* dz, const float* w, int ci_off, int ci_n, const float* y_prev,
                         int act_prev, float* dx, void* workspace, long long workspace_bytes, void* stream) {
  if (int rc = check_desc(d, "conv_dgrad_ex")) return rc;
  if (d->in_mode != CLSLAM_IN_PLAIN && ci_off < d->c1) { set_error("conv_dgrad_ex: an UPCAT layer's up-sampled part goes through conv_dgrad"); return -1; }
  if (d->stride != 1 && d->pad_mode != CLSLAM_PAD_ZERO) { set_error("conv_dgrad_ex: strided layers are zero padded"); return -1; }
  if (ci_off < 0 || ci_n < 4 || ci_off % 4 || ci_n % 4 || ci_off + ci_n > d->ci) { set_error("conv_dgrad_ex: bad channel slice"); return -1; }
  DgradIn g;
  g.dz = dz; g.n = d->n; g.h = d->ho; g.w = d->wo; g.co = d->co;
  g.kh = d->kh; g.kw = d->kw; g.pad = d->pad; g.pad_mode = d->pad_mode;
  g.up = 0; g.stride = d->stride; g.hl = d->hi; g.wl = d->wi;          // the slice lives at the layer's input resolution
  const int M = d->n * d->hi * d->wi, N = ci_n, K = d->kh * d->kw * d->co;
  ADgrad a{g, M, K};
  BWeightDgradOp b{{w + ci_off, N, d->ci, d->co, d->kh * d->kw}};
  EpiDgrad e{dx, y_prev, act_prev, nullptr, nullptr, N, g.hl, g.wl};
  return run_gemm(a, b, e, M, N, K, (float*)workspace, workspace_bytes, as_stream(stream), "conv_dgrad_ex");
}

long long clslam_conv_dgrad_ex_workspace_bytes(const clslam_conv_desc* d, int ci_n) {
  if (check_desc(d, "conv_dgrad_ex_ws")) return -1;
  return ws_need(d->n * d->hi * d->wi, ci_n, d->kh * d->kw * d->co, false);
}

long long clslam_conv_wgrad_workspace_bytes(const clslam_conv_desc* d) {
  if (check_desc(d, "conv_wgrad_ws")) return -1;
  const int M = d->co, N = d->kh * d->kw * d->ci, K = d->n * d->ho * d->wo;
  return ws_need(M, N, K, true);
}

int clslam_conv_wgrad(const clslam_conv_desc* d, const float* x, const float* x2, const float* dz, float* dw, float* db,
                      void* workspace, long long workspace_bytes, void* stream) {
  if (int rc = check_desc(d, "conv_wgrad")) return rc;
  const int M = d->co, N = d->kh * d->kw * d->ci, K = d->n * d->ho * d->wo;
  AWgradDz a{dz, d->co, K};
  BWgradIn b{make_conv_in(*d, x, x2), N};
  if (g_conv_mode == 1 && K >= 32) {
    EpiWgradTc e{dw, db, N};
    return launch_gemm_tc(plan_tc(M, N + 4, K), a, b, e, M, N + 4, K, (float*)workspace, workspace_bytes, as_stream(stream),
                          "conv_wgrad");
  }
  EpiWgrad e{dw, db, N};
  return run_gemm(a, b, e, M, N, K, (float*)workspace, workspace_bytes, as_stream(stream), "conv_wgrad");
}

}  // extern "C"
