// TMA-fed tcgen05 weight-gradient kernel for 3x3 / stride-1 convolutions over bf16 hi/lo planes
// (include/clslam_b200.h: clslam_wgrad3x3_tma).
//
//   dW[co][r][s][ci] = sum over pixels p of dz[p][co] * xpad[p + (r, s)][ci],   db[co] = sum_p dz[p][co]
//
// GEMM with K = pixels.  Both operands are channel-contiguous (NHWC), i.e. "MN-major" for the tensor core, which
// tcgen05 reads directly through MN-major shared-memory descriptors -- no transposition anywhere:
//   A (M side) = the layer input: for one column shift s a patch of (R + 2) rows x 16 pixels x CC channels; the
//       128 rows of the MMA are blocks of CC channels taken at patch rows rr + 0, rr + 1, rr + 2, ... (leading-dimension
//       byte offset = one patch row), so ONE instruction covers the three vertical taps r = 0..2 (the remaining
//       blocks read rows further down: finite junk that lands in accumulator lanes nobody reads);
//   B (N side) = dz: 16 pixels of tile row rr x CO channels;
//   D_s[(r, ci)][co] (fp32 in TMEM, one accumulator per horizontal tap s) accumulates over ALL tiles of the CTA.
// Per 16 pixels and s: A_lo.B_hi + A_hi.B_lo + A_hi.B_hi.  Each CTA owns one (co block, ci chunk) pair and a share of
// the pixel tiles; partial results go to the workspace and a second kernel adds them in a fixed order
// (deterministic) into dW [co][3][3][ci].  The bias gradient db[co] = sum_p dz[p][co] is a separate two-stage
// channel reduction over the fp32 dz (bias_grad kernels below).
//
//   warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: final TMEM -> workspace copy.
#include <cuda.h>
#include <cstdlib>
#include <cudaTypedefs.h>

#include "planes.cuh"

namespace clslam {

constexpr int kWgR = 4;                    // tile rows
constexpr int kWgW = 16;                   // tile columns = K of one MMA
constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 8;
constexpr int kWgSmemMax = 232448;

template <int CC, int CO>
struct WgCfg {
  static constexpr int kRowA = CC * 2;                        // bytes of one pixel of the x patch
  static constexpr int kAtomCo = CO > 64 ? 64 : CO;           // channels per dz box (<= 128 bytes: the swizzle span)
  static constexpr int kAtoms = CO / kAtomCo;
  static constexpr int kRowB = kAtomCo * 2;
  static constexpr int kAPlane = (kWgR + 2) * kWgW * kRowA;
  static constexpr int kAShift = 2 * kAPlane;                 // hi + lo of one column shift
  static constexpr int kABytes = 3 * kAShift;
  static constexpr int kBPlane = kWgR * kWgW * kRowB;
  static constexpr int kBAtom = 2 * kBPlane;
  static constexpr int kBBytes = kAtoms * kBAtom;
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kTail = 8 * kWgW * kRowA;              // the junk row blocks of the last stage stay inside the allocation
  static constexpr int kCols = 3 * CO;                        // workspace columns: one block per horizontal tap
  // N-concatenation (CO <= 64: one channel atom per plane): x_hi . [dz_hi | dz_lo] in one MMA (the lo plane follows the hi
  // plane, leading-dimension offset = one plane) + x_lo . dz_hi: two instructions per 16 pixels and tap instead of three
  static constexpr bool kNcat = CO <= 64;
  static constexpr int kAccCols = kNcat ? 2 * CO : CO;        // TMEM columns per horizontal tap
  static constexpr int kTmemNeed = 3 * kAccCols;
  static constexpr int kTmemCols = kTmemNeed <= 32 ? 32 : (kTmemNeed <= 64 ? 64 : (kTmemNeed <= 128 ? 128 : (kTmemNeed <= 256 ? 256 : 512)));
  static constexpr uint32_t layout_of(int row_bytes) { return row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u); }
  // upper descriptor words: SBO (stride between 8-pixel groups) | version 1 | swizzle mode
  static constexpr uint32_t kDescHiA = (uint32_t)((8 * kRowA) >> 4) | (1u << 14) | (layout_of(kRowA) << 29);
  static constexpr uint32_t kDescHiB = (uint32_t)((8 * kRowB) >> 4) | (1u << 14) | (layout_of(kRowB) << 29);
  static constexpr uint32_t kLboA = (uint32_t)((kWgW * kRowA) >> 4) << 16;       // next row block = next patch row
  static constexpr uint32_t kLboB = (uint32_t)((kNcat ? kBPlane : kBAtom) >> 4) << 16;   // next N block: the lo plane (N-concat) or the next 64-channel atom
  static_assert(kStage % 1024 == 0, "stage alignment");
  static_assert(kTmemNeed <= 512, "TMEM");
};

struct WgArgs {
  int n_img, h, w;                 // output (= dz) grid
  int tiles_x, tiles_y;
  int co_blocks, ci_chunks, splits;
  int org_x;                       // x-patch origin relative to the output pixel in x-map coordinates (0: padded view, -1: interior)
  int stages;
  FastDiv div_tpi, div_tx;         // divisions by tiles per image and tiles_x
  float* ws;                       // [cta][3 * CC][3 * CO]
};

// instruction descriptor: D f32, A/B bf16, both MN-major, M = 128, N = n
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ uint64_t wg_desc(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

template <int CC, int CO>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad3x3_tma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz, const WgArgs g) {
  using C = WgCfg<CC, CO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)g.stages * C::kStage + C::kTail);
  uint64_t* empty = full + kWgMaxStages;
  uint64_t* done = empty + kWgMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgMaxStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_dz); }
  if (warp == 2) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();            // the prologue above overlapped the previous kernel; its results are visible from here on
  pdl_trigger();

  // work item of this CTA: (co block, ci chunk, split of the pixel tiles)
  const int pair = blockIdx.x / g.splits, split = blockIdx.x - pair * g.splits;
  const int cb = pair / g.ci_chunks, chunk = pair - cb * g.ci_chunks;
  const int tiles_per_img = g.tiles_x * g.tiles_y;
  const int total = g.n_img * tiles_per_img;
  const int t_begin = (int)(((long long)total * split) / g.splits), t_end = (int)(((long long)total * (split + 1)) / g.splits);

  if (warp == 0) {
    // TMA producer: warp-uniform loop, one elected lane issues
    int st = 0, ph = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int img = g.div_tpi.quot(t), rem = t - img * tiles_per_img;
      const int ty = g.div_tx.quot(rem);
      const int h0 = ty * kWgR, w0 = (rem - ty * g.tiles_x) * kWgW;
      mbar_wait(&empty[st], (uint32_t)(ph ^ 1));
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[st], C::kStage);
        uint8_t* base = tiles + (size_t)st * C::kStage;
#pragma unroll
        for (int s = 0; s < 3; ++s)
          tma_load_5d(base + s * C::kAShift, &map_x, &full[st], chunk * CC, w0 + s + g.org_x, h0 + g.org_x, img, 0);
#pragma unroll
        for (int a = 0; a < C::kAtoms; ++a)
          tma_load_5d(base + C::kABytes + a * C::kBAtom, &map_dz, &full[st], cb * CO + a * C::kAtomCo, w0, h0, img, 0);
      }
      __syncwarp();
      if (++st == g.stages) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop (descriptors in uniform registers), one elected lane issues
    constexpr uint32_t idesc = umma_idesc_bf16_mn(CO);
    constexpr uint32_t idesc2 = umma_idesc_bf16_mn(2 * CO);
    const uint32_t ring = (smem_u32(tiles) >> 4) & 0x3FFF;
    int st = 0, ph = 0;
    uint32_t acc = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&full[st], (uint32_t)ph);
      tc_fence_after();
      const uint32_t a0 = ring + (uint32_t)(st * (C::kStage >> 4));
      const uint32_t b0 = a0 + (uint32_t)(C::kABytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int rr = 0; rr < kWgR; ++rr) {
          const uint32_t b_hi = (b0 + (uint32_t)((rr * kWgW * C::kRowB) >> 4)) | C::kLboB;
          const uint32_t b_lo = b_hi + (uint32_t)(C::kBPlane >> 4);
          const uint64_t dbh = wg_desc(C::kDescHiB, b_hi), dbl = wg_desc(C::kDescHiB, b_lo);
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const uint32_t a_hi = (a0 + (uint32_t)((s * C::kAShift + rr * kWgW * C::kRowA) >> 4)) | C::kLboA;
            const uint32_t a_lo = a_hi + (uint32_t)(C::kAPlane >> 4);
            const uint64_t dah = wg_desc(C::kDescHiA, a_hi), dal = wg_desc(C::kDescHiA, a_lo);
            const uint32_t d = tmem_base + (uint32_t)(s * C::kAccCols);
            if constexpr (C::kNcat) {
              umma_bf16(d, dah, dbh, idesc2, rr == 0 ? acc : 1u);       // x_hi . [dz_hi | dz_lo]
              umma_bf16(d, dal, dbh, idesc, 1u);                        // + x_lo . dz_hi
            } else {
              umma_bf16(d, dal, dbh, idesc, rr == 0 ? acc : 1u);
              umma_bf16(d, dah, dbl, idesc, 1u);
              umma_bf16(d, dah, dbh, idesc, 1u);
            }
          }
        }
        umma_commit(&empty[st]);
      }
      __syncwarp();
      acc = 1u;
      if (++st == g.stages) { st = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
  } else {
    // final copy: lanes (r, ci) < 3 * CC, columns (s, co)
    if (lane == 0) mbar_wait(done, 0);
    __syncwarp();
    tc_fence_after();
    const int q = warp & 3;
    const int l = q * 32 + lane;
    float* dst = g.ws + ((size_t)blockIdx.x * (3 * CC) + l) * C::kCols;
#pragma unroll 1
    for (int c = 0; c < C::kCols; c += 8) {
      uint32_t r[8];
      if constexpr (C::kNcat) {             // tap block s = c / CO: its two column halves (x_hi.dz_hi + x_lo.dz_hi, x_hi.dz_lo) are added here
        const int sblk = c / CO, cc = c - sblk * CO;
        uint32_t r2[8];
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sblk * C::kAccCols + cc), r);
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sblk * C::kAccCols + CO + cc), r2);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
      } else {
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      }
      if (l < 3 * CC && t_end > t_begin) {
        *reinterpret_cast<float4*>(dst + c) = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
        *reinterpret_cast<float4*>(dst + c + 4) = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
      } else if (l < 3 * CC) {
        *reinterpret_cast<float4*>(dst + c) = f4zero();
        *reinterpret_cast<float4*>(dst + c + 4) = f4zero();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
}

// dW[co][r][s][ci] = sum over splits of ws[(cb, chunk, split)][r * CC + ci % CC][s * CO + co % CO]
// ... and db[o] = sum over the bias_grad_partial blocks (the blocks past the dW tiles)
// One block = one tap x 32 output channels x 8 input channels x G groups of partial sums (blockDim = 32 x 8 x G): the partial
// sums are read along co (contiguous in the workspace), every thread adds its group's share of the `splits` CTA partials with
// eight independent chains (the kernel is a latency-bound column sum of a [splits][9 ci co] matrix: what counts is loads in
// flight), the G group sums are added in a fixed order (deterministic), and the result is transposed through shared memory and
// written along ci (contiguous in dW, one 32-byte sector per 8 threads).  The round-2a kernel used 32 x 32 tiles with one
// thread per 4 elements and no groups: 9 .. 36 blocks for the thin decoder layers, 25 - 35 us per launch at 0.6 TB/s.
__global__ void __launch_bounds__(1024)
wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int co, int ci, int cc, int cob, int ci_chunks, int splits,
                    const float* __restrict__ bias_part, int bias_blocks, float* __restrict__ db, int tiles_o, int tiles_c) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[4][8][33];
  const int tx = threadIdx.x, ty = threadIdx.y, gz = threadIdx.z, G = blockDim.z;
  const int tid = ty * 32 + tx;                 // within the group
  const int n_tiles = 9 * tiles_o * tiles_c;
  if ((int)blockIdx.x >= n_tiles) {          // bias blocks: fixed-order sum of the per-block channel sums
    const int o = ((int)blockIdx.x - n_tiles) * 256 + tid;
    if (gz == 0 && db && o < co) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int b = 0;
      for (; b + 4 <= bias_blocks; b += 4) {
        a0 += bias_part[(size_t)b * co + o]; a1 += bias_part[(size_t)(b + 1) * co + o];
        a2 += bias_part[(size_t)(b + 2) * co + o]; a3 += bias_part[(size_t)(b + 3) * co + o];
      }
      for (; b < bias_blocks; ++b) a0 += bias_part[(size_t)b * co + o];
      db[o] = (a0 + a1) + (a2 + a3);
    }
    return;
  }
  const int cols = 3 * cob;
  const size_t cta_stride = (size_t)(3 * cc) * cols;
  const int t = blockIdx.x;
  const int tc = t % tiles_c, to = (t / tiles_c) % tiles_o, tap = t / (tiles_c * tiles_o);
  const int r = tap / 3, s = tap - r * 3;
  const int o = to * 32 + tx, c = tc * 8 + ty;
  float v = 0.f;
  if (o < co && c < ci) {
    const int pair = (o / cob) * ci_chunks + c / cc;
    const float* p = ws + (size_t)pair * splits * cta_stride + (size_t)(r * cc + c % cc) * cols + s * cob + o % cob;
    const int k0 = (splits * gz) / G, k1 = (splits * (gz + 1)) / G;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f;
    int k = k0;
    for (; k + 8 <= k1; k += 8) {
      a0 += p[(size_t)k * cta_stride]; a1 += p[(size_t)(k + 1) * cta_stride];
      a2 += p[(size_t)(k + 2) * cta_stride]; a3 += p[(size_t)(k + 3) * cta_stride];
      a4 += p[(size_t)(k + 4) * cta_stride]; a5 += p[(size_t)(k + 5) * cta_stride];
      a6 += p[(size_t)(k + 6) * cta_stride]; a7 += p[(size_t)(k + 7) * cta_stride];
    }
    for (; k + 2 <= k1; k += 2) { a0 += p[(size_t)k * cta_stride]; a1 += p[(size_t)(k + 1) * cta_stride]; }
    if (k < k1) a2 += p[(size_t)k * cta_stride];
    v = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  }
  tile[gz][ty][tx] = v;
  __syncthreads();
  if (gz == 0) {
    const int wc = tid & 7, wo = tid >> 3;          // 8 consecutive threads = 8 consecutive input channels of one (co, tap)
    const int oo = to * 32 + wo, cw = tc * 8 + wc;
    if (oo < co && cw < ci) {
      float acc = tile[0][wc][wo];
      for (int g2 = 1; g2 < G; ++g2) acc += tile[g2][wc][wo];
      dw[((size_t)oo * 9 + tap) * ci + cw] = acc;
    }
  }
}

// db[c] = sum over pixels of dz[p][c] (fp32 NHWC), deterministic: per-block partial sums, then a fixed-order reduction
constexpr int kBiasBlocks = 592;          // 4 per SM: the kernel is a latency-bound streaming sum
__global__ void __launch_bounds__(256)
bias_grad_partial_kernel(const float* __restrict__ dz, long long pixels, int c, float* __restrict__ part) {
  pdl_wait();
  pdl_trigger();
  // thread = (pixel lane, channel quad); c / 4 quads (<= 64), 256 / quads pixel lanes
  __shared__ float4 red[256];
  const int quads = c / 4, lanes = 256 / quads;
  const int qd = threadIdx.x % quads, pl = threadIdx.x / quads;
  const long long per = (pixels + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per, p1 = p0 + per < pixels ? p0 + per : pixels;
  float4 acc = f4zero();
  if (pl < lanes) {
    // four independent loads in flight per thread; the four chains are combined in a fixed order (deterministic)
    float4 a0 = f4zero(), a1 = f4zero(), a2 = f4zero(), a3 = f4zero();
    long long p = p0 + pl;
    for (; p + 3LL * lanes < p1; p += 4LL * lanes) {
      const float4 v0 = ld4(dz + p * c + qd * 4), v1 = ld4(dz + (p + lanes) * c + qd * 4);
      const float4 v2 = ld4(dz + (p + 2LL * lanes) * c + qd * 4), v3 = ld4(dz + (p + 3LL * lanes) * c + qd * 4);
      f4add(a0, v0); f4add(a1, v1); f4add(a2, v2); f4add(a3, v3);
    }
    for (; p < p1; p += lanes) f4add(a0, ld4(dz + p * c + qd * 4));
    f4add(a0, a1); f4add(a2, a3); f4add(a0, a2);
    acc = a0;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < quads) {
    float4 s = f4zero();
    for (int k = 0; k < lanes; ++k) f4add(s, red[k * quads + threadIdx.x]);
    *reinterpret_cast<float4*>(part + (size_t)blockIdx.x * c + threadIdx.x * 4) = s;
  }
}
// ---- host side ---------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled g_wg_encode = nullptr;
static int wg_get_encoder() {
  if (g_wg_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) { set_error("cuTensorMapEncodeTiled is not available"); return -3; }
  g_wg_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  return 0;
}

static int wg_encode(CUtensorMap* m, const clslam_planes& p, bool padded_view, int box_c, int box_rows, const char* what) {
  const int hp = p.h + 2, wp = p.w + 2;
  __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(p.base) + p.c_off + (padded_view ? 0 : ((size_t)wp + 1) * p.c_total);
  const cuuint64_t dims[5] = {(cuuint64_t)p.c, (cuuint64_t)(padded_view ? wp : p.w), (cuuint64_t)(padded_view ? hp : p.h), (cuuint64_t)p.n, 2};
  const cuuint64_t str[4] = {(cuuint64_t)p.c_total * 2, (cuuint64_t)wp * p.c_total * 2, (cuuint64_t)hp * wp * p.c_total * 2,
                             (cuuint64_t)p.plane_stride * 2};
  const cuuint32_t box[5] = {(cuuint32_t)box_c, kWgW, (cuuint32_t)box_rows, 1, 2};
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  const int row_bytes = box_c * 2;
  const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const CUresult r = g_wg_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("wgrad3x3_tma: cuTensorMapEncodeTiled(%s) failed: %d", what, (int)r); return -4; }
  return 0;
}

struct WgPlan { int cc, cob, co_blocks, ci_chunks, splits, grid, stages, smem; long long ws_bytes, partial_bytes; };

static int wg_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    // CLSLAM_WGRAD_SMS: CTA budget of a weight-gradient launch (default: every SM).  The launches run on a side lane next to the
    // input-gradient chain; fewer CTAs leave SMs to that chain and mean fewer partial sums for wgrad_reduce_kernel
    if (const char* e = getenv("CLSLAM_WGRAD_SMS")) { const int v = atoi(e); if (v >= 8 && v < sms) sms = v; }
  }
  return sms;
}

template <int CC, int CO>
static void wg_fill(WgPlan& p) {
  using C = WgCfg<CC, CO>;
  const int fixed = 1024 + C::kTail + 256;
  int st = (kWgSmemMax - fixed) / C::kStage;
  p.stages = st > kWgMaxStages ? kWgMaxStages : st;
  p.smem = p.stages * C::kStage + fixed;
}

static int wg_plan(int n, int h, int w, int ci, int co, WgPlan& p) {
  if (ci % 16 || co % 16) return -1;
  p.cc = ci % 32 == 0 ? 32 : 16;
  p.cob = co >= 128 ? 128 : co;
  if (co % p.cob || (p.cob != 16 && p.cob != 32 && p.cob != 64 && p.cob != 128)) return -1;
  if (p.cc == 16 && p.cob != 16) return -1;          // instantiated: (32, 16 / 32 / 64 / 128) and (16, 16)
  p.co_blocks = co / p.cob;
  p.ci_chunks = ci / p.cc;
  const int pairs = p.co_blocks * p.ci_chunks;
  const int tiles = n * ((h + kWgR - 1) / kWgR) * ((w + kWgW - 1) / kWgW);
  int splits = wg_sms() / pairs;
  if (splits < 1) splits = 1;
  if (splits > tiles) splits = tiles;
  p.splits = splits;
  p.grid = pairs * splits;
  if (p.cc == 16) wg_fill<16, 16>(p);
  else if (p.cob == 16) wg_fill<32, 16>(p);
  else if (p.cob == 32) wg_fill<32, 32>(p);
  else if (p.cob == 64) wg_fill<32, 64>(p);
  else wg_fill<32, 128>(p);
  p.partial_bytes = ((long long)p.grid * (3 * p.cc) * (3 * p.cob) * 4 + 255) / 256 * 256;
  p.ws_bytes = p.partial_bytes + (long long)kBiasBlocks * co * 4;        // + the bias_grad partial sums
  return p.stages >= 2 ? 0 : -1;
}

template <int CC, int CO>
static int wg_launch(const CUtensorMap& mx, const CUtensorMap& mdz, const WgArgs& g, const WgPlan& p, cudaStream_t st) {
  auto kern = wgrad3x3_tma_kernel<CC, CO>;
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemMax);
    if (e != cudaSuccess) { set_error("wgrad3x3_tma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -3; }
    attr_set = true;
  }
  launch_k(kern, dim3(p.grid), dim3(kWgThreads), p.smem, st, mx, mdz, g);
  return finish_launch("wgrad3x3_tma");
}

}  // namespace clslam

using namespace clslam;

extern "C" {

long long clslam_wgrad3x3_tma_workspace_bytes(int n, int h, int w, int ci, int co) {
  WgPlan p{};
  if (wg_plan(n, h, w, ci, co, p)) { set_error("wgrad3x3_tma: unsupported channels (ci=%d co=%d)", ci, co); return -1; }
  return p.ws_bytes;
}

int clslam_wgrad3x3_tma(const clslam_planes* x, int pad_mode, const clslam_planes* dz, const float* dz_f32, float* dw, float* db,
                        void* workspace, long long workspace_bytes, void* stream) {
  if (!x || !dz || !x->base || !dz->base || !dw) { set_error("wgrad3x3_tma: null argument"); return -1; }
  if (pad_mode != CLSLAM_PAD_ZERO && pad_mode != CLSLAM_PAD_REFLECT) { set_error("wgrad3x3_tma: pad_mode"); return -1; }
  if (x->n != dz->n || x->h != dz->h || x->w != dz->w) { set_error("wgrad3x3_tma: x / dz dims differ"); return -1; }
  if (x->c_total % 8 || x->c_off % 8 || dz->c_total % 8 || dz->c_off % 8) { set_error("wgrad3x3_tma: channel alignment"); return -1; }
  WgPlan p{};
  if (wg_plan(x->n, x->h, x->w, x->c, dz->c, p)) { set_error("wgrad3x3_tma: unsupported channels (ci=%d co=%d)", x->c, dz->c); return -1; }
  if (!workspace || workspace_bytes < p.ws_bytes) { set_error("wgrad3x3_tma: workspace too small (%lld < %lld)", workspace_bytes, p.ws_bytes); return -2; }
  if (int rc = wg_get_encoder()) return rc;
  CUtensorMap mx, mdz;
  const bool padded = pad_mode == CLSLAM_PAD_REFLECT;
  if (int rc = wg_encode(&mx, *x, padded, p.cc, kWgR + 2, "x")) return rc;
  if (int rc = wg_encode(&mdz, *dz, false, p.cob > 64 ? 64 : p.cob, kWgR, "dz")) return rc;
  WgArgs g{};
  g.n_img = x->n; g.h = x->h; g.w = x->w;
  g.tiles_x = (x->w + kWgW - 1) / kWgW; g.tiles_y = (x->h + kWgR - 1) / kWgR;
  g.co_blocks = p.co_blocks; g.ci_chunks = p.ci_chunks; g.splits = p.splits;
  g.org_x = padded ? 0 : -1;
  g.div_tpi = make_fastdiv(g.tiles_x * g.tiles_y); g.div_tx = make_fastdiv(g.tiles_x);
  g.stages = p.stages;
  g.ws = reinterpret_cast<float*>(workspace);
  cudaStream_t st = as_stream(stream);
  int rc;
  if (p.cc == 16) rc = wg_launch<16, 16>(mx, mdz, g, p, st);
  else if (p.cob == 16) rc = wg_launch<32, 16>(mx, mdz, g, p, st);
  else if (p.cob == 32) rc = wg_launch<32, 32>(mx, mdz, g, p, st);
  else if (p.cob == 64) rc = wg_launch<32, 64>(mx, mdz, g, p, st);
  else rc = wg_launch<32, 128>(mx, mdz, g, p, st);
  if (rc) return rc;
  float* bias_part = nullptr;
  if (db) {       // bias gradient from the fp32 dz: per-block channel sums, folded into the reduction kernel below
    if (!dz_f32 || dz->c > 256) { set_error("wgrad3x3_tma: db needs dz_f32 and co <= 256"); return -1; }
    bias_part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + p.partial_bytes);
    const long long pixels = (long long)dz->n * dz->h * dz->w;
    launch_k(bias_grad_partial_kernel, dim3(kBiasBlocks), dim3(256), 0, st, dz_f32, pixels, dz->c, bias_part);
    if ((rc = finish_launch("bias_grad"))) return rc;
  }
  const int tiles_o = (dz->c + 31) / 32, tiles_c = (x->c + 7) / 8;
  const int blocks = 9 * tiles_o * tiles_c + (db ? (dz->c + 255) / 256 : 0);
  const int groups = p.splits >= 32 ? 4 : (p.splits >= 16 ? 2 : 1);          // partial-sum groups per element (loads in flight)
  launch_k(wgrad_reduce_kernel, dim3(blocks), dim3(32, 8, groups), 0, st, g.ws, dw, dz->c, x->c, p.cc, p.cob, p.ci_chunks, p.splits,
           (const float*)bias_part, (int)kBiasBlocks, db, tiles_o, tiles_c);
  return finish_launch("wgrad_reduce");
}

}  // extern "C"
