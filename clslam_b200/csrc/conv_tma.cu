// TMA-fed, persistent, warp-specialised tcgen05 kernel for 3x3 / stride-1 convolutions over bf16
// hi/lo activation planes (include/clslam_b200.h: clslam_conv3x3_tma).
//
//   warp 0    : TMA producer   (one elected lane issues cp.async.bulk.tensor loads)
//   warp 1    : MMA issuer     (one elected lane issues tcgen05.mma / tcgen05.commit)
//   warps 2-17: epilogue       (tcgen05.ld -> scale/shift/residual/activation -> fp32 and/or planes)
//
// M tile = 16 output rows x 8 output columns of one image (TMEM lane l <-> pixel (l / 8, l % 8)): an
// 8-pixel row segment is one 8-row core-matrix group of the K-major swizzled UMMA layout, so the input
// patch loaded for column shift s (box: KC channels x 8 px x 18 rows, both planes in one TMA
// instruction) serves the three taps (r, s), r = 0..2, by moving the descriptor start by r groups.
// Per (channel chunk, s) one A stage; per (chunk, s, r) one B stage (weights [BN][KC], hi + lo).
// Per k16 step three MMAs: A_lo.B_hi + A_hi.B_lo + A_hi.B_hi (fp32 accumulate in TMEM, two accumulator
// stages so the epilogue of tile i overlaps the main loop of tile i + 1).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>

#include "conv_ops.cuh"
#include "planes.cuh"

namespace clslam {

constexpr int kTmaTH = 16, kTmaTW = 8;     // output tile
// timing experiments (clslam_debug_tma_flags) are compiled out unless CLSLAM_TMA_DEBUG is defined: their flag tests sit in
// the single-thread issue loops, where every instruction counts
#ifdef CLSLAM_TMA_DEBUG
#define TMA_DBG(g, bit) (((g).debug & (bit)) != 0)
// clock64 stamps of CTA 0's first work item (read back with clslam_debug_tma_stamps; tools/tma_stamps.py names them)
__device__ unsigned long long g_tma_stamps[32];
#define TMA_STAMP(i, cond) do { if (blockIdx.x == 0 && (cond)) g_tma_stamps[i] = clock64(); } while (0)
#else
#define TMA_DBG(g, bit) (false)
#define TMA_STAMP(i, cond) do { } while (0)
#endif
constexpr int kTmaEpiWarps = 16;          // 4 TMEM lane quarters x 4 column slots (8 accumulator columns per tcgen05.ld)
constexpr int kTmaThreads = 64 + 32 * kTmaEpiWarps;

// ELU on the epilogue's hot path: exp through ex2.approx, Taylor series next to 0 where exp(x) - 1 cancels
// (relative error < 3e-6 of the result; the reference's expm1f costs ~3x the instructions)
__device__ __forceinline__ float elu_fast(float v) {
  const float e = __expf(v) - 1.f;
  const float t = v * (1.f + v * (0.5f + v * (0.16666667f + v * (0.041666668f + v * 0.0083333338f))));
  const float neg = v > -0.0625f ? t : e;
  return v > 0.f ? v : neg;
}
__device__ __forceinline__ float act_fast(float v, int act) {
  if (act == CLSLAM_ACT_ELU) return elu_fast(v);
  return act_apply(v, act);
}

template <int KC, int BN, int T = 3>      // T x T taps (3: the 3x3 layers; 4: conv1 in its space-to-depth form)
struct TmaConvCfg {
  static constexpr int kRowBytes = KC * 2;
  static constexpr int kPH = kTmaTH + T - 1;                  // rows of an A patch
  static constexpr int kAPlane = kPH * kTmaTW * kRowBytes;
  static constexpr int kAStage = 2 * kAPlane;
  static constexpr int kBPlane = BN * kRowBytes;
  static constexpr int kBStage = 2 * kBPlane;
  static constexpr int kSlots = BN >= 32 ? 4 : BN / 8;        // epilogue column slots with work
  static constexpr int kAcc = BN <= 64 ? 4 : 2;               // accumulator stages in TMEM (thin layers: the hand-off latency dominates)
  // N-concatenation: with BN <= 64 one MMA multiplies A_hi by [B_hi; B_lo] (the lo plane follows the hi plane in the
  // weight tile, N = 2 BN) into two column blocks that the epilogue adds, a second MMA adds A_lo.B_hi: two instructions
  // per k step instead of three (an SS-mode MMA with M = 128 costs ~64 cycles for any N <= 128: it is bound by the A read)
  static constexpr bool kNcat = BN <= 64;
  static constexpr int kAccCols = kNcat ? 2 * BN : BN;        // TMEM columns of one accumulator stage
  static constexpr int kTmemCols = kAcc * kAccCols < 32 ? 32 : kAcc * kAccCols;
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns");
  // upper 32 bits of every K-major swizzled descriptor of this configuration (SBO = 8 rows, version 1, layout type)
  static constexpr uint32_t kDescHi = (uint32_t)((8 * kRowBytes) >> 4) | (1u << 14) |
                                      ((kRowBytes == 128 ? 2u : (kRowBytes == 64 ? 4u : 6u)) << 29);
  static_assert(kAStage % (8 * kRowBytes) == 0 && kBStage % (8 * kRowBytes) == 0, "stage alignment (swizzle atom)");
};
constexpr int kTmaMaxStages = 8;
constexpr int kTmaBarBytes = 512;          // 4 x 8 ring barriers + 2 x 4 accumulator barriers + resident-B barrier + TMEM slot
constexpr int kTmaSmemMax = 232448;
constexpr int kTmaResidentMax = 128 * 1024;

struct TmaConvArgs {
  int n_img, ho, wo, cout, cin;
  int tiles_x, tiles_y, n_blocks, chunks;
  int org;                       // input coordinate of tap (0, 0) relative to the output pixel, in tensor-map coordinates
  FastDiv div_nb, div_tpi, div_tx;   // divisions by n_blocks, tiles per image, tiles_x
  int debug;                     // timing experiments only (results are wrong when != 0): 1 no MMA, 2 no epilogue body, 4 no B ring, 8 no A loads
  int na, nbr;                   // A ring depth; B ring depth, or 0: the whole weight tensor stays resident in shared memory
  int act;
  const float* scale;
  const float* shift;
  PlaneView res;                 // base == nullptr: none
  const float* gate;             // optional fp32 NHWC (n, ho, wo, cout): result *= act'(gate) (dgrad through the previous activation)
  int gate_act;
  float* y;
  PlaneView out;                 // base == nullptr: none
  // split-K form (SK instantiations only; appended so that the offsets above are those of the plain kernel): work item
  // t = (tile, N block) * ksplit + z handles the (channel chunk, column shift) steps [z * cs_per, (z + 1) * cs_per) and
  // stores its raw fp32 accumulators to part[z][pixel][cout]; tma_splitk_epilogue_kernel adds them in a fixed order
  int ksplit, cs_per;
  FastDiv div_ks;
  float* part;
  long long part_stride;         // elements between two z slices = n_img * ho * wo * cout
  // fused split-K reduction: one counter per (tile, N block).  The CTA whose atomic increment finds ksplit - 1 is the last
  // slice of that tile to land: its epilogue warps add the ksplit partial sums (fixed z order: the result does not depend on
  // which CTA came last) and run the real epilogue, then reset the counter for the next launch.  No second kernel, nobody
  // spins.  nullptr: partial sums only (tma_splitk_epilogue_kernel / dgrad_fold_splitk_kernel reduce them).
  int* counters;
  // cluster split-K: the ksplit slices of a tile are the CTAs of ONE thread-block cluster (launch attribute); every CTA stages
  // its fp32 accumulator tile in its own shared memory, and after a cluster barrier CTA r adds rows [r, r + 1) * 128 / ksplit
  // of all ksplit tiles over distributed shared memory (fixed rank order: deterministic) and runs the epilogue for them.  The
  // partial sums never leave the chip, there is no second kernel and the epilogue is spread over the cluster.  0 = off
  int cluster_ks;
  // tall form (TALL instantiations only): the input is addressed as ONE image of tall_n * tall_hp rows -- the images of the
  // planes buffer stacked with their halo rows, which are valid padding (reflection halo, or zeros for a zero-padded layer
  // whose producer left the halo untouched).  A 16-row tile then covers 16 / tall_hp images instead of wasting 16 - h rows:
  // 6 x 20 maps (tall_hp = 8) need half the tiles.  n_img = 1, ho stays the height of one image, output row v of the
  // stack is row v % tall_hp of image v / tall_hp (valid when < ho)
  int tall_hp, tall_n;
  FastDiv div_hp;
  int grid_cap;                  // host side only: clslam_conv_tma.sm_budget (0 = every SM)
};


// ---- cluster split-K helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem4(uint32_t local_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
  return v;
}

// the epilogue proper for 8 output channels [n, n + 8) of output pixel (img, oh, ow): the same sequence as the fused epilogue
// of conv3x3_tma_kernel (scale, shift, residual, activation, gate; fp32 and / or planes with nearest-2x footprint and halo)
__device__ __forceinline__ void tma_epilogue_store(const TmaConvArgs& g, int img, int oh, int ow, int n, float (&v)[8]) {
  const size_t y_off = (((size_t)img * g.ho + oh) * g.wo + ow) * g.cout;
  if (g.scale) {
    const float4 s0 = ld4(g.scale + n), s1 = ld4(g.scale + n + 4);
    v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w; v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
  }
  if (g.shift) {
    const float4 s0 = ld4(g.shift + n), s1 = ld4(g.shift + n + 4);
    v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w; v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
  }
  if (g.res.base) {
    float r[8];
    planes_load8(g.res, img, oh + 1, ow + 1, n, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += r[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = act_fast(v[j], g.act);
  if (g.gate) {
    const float4 g0 = ld4(g.gate + y_off + n), g1 = ld4(g.gate + y_off + n + 4);
    v[0] *= act_grad_from_output(g0.x, g.gate_act); v[1] *= act_grad_from_output(g0.y, g.gate_act);
    v[2] *= act_grad_from_output(g0.z, g.gate_act); v[3] *= act_grad_from_output(g0.w, g.gate_act);
    v[4] *= act_grad_from_output(g1.x, g.gate_act); v[5] *= act_grad_from_output(g1.y, g.gate_act);
    v[6] *= act_grad_from_output(g1.z, g.gate_act); v[7] *= act_grad_from_output(g1.w, g.gate_act);
  }
  if (g.y) {
    float* p = g.y + y_off + n;
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (g.out.base) planes_emit<8>(g.out, img, oh, ow, n, v);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

template <int KC, int BN, int T, bool SK = false, bool TALL = false>
__global__ void __launch_bounds__(kTmaThreads, 1)
conv3x3_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TmaConvArgs g) {
  using C = TmaConvCfg<KC, BN, T>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* a_tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* b_tiles = a_tiles + g.na * C::kAStage;
  const int b_slots = g.nbr > 0 ? g.nbr : g.chunks * T * T;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(b_tiles + (size_t)b_slots * C::kBStage);
  uint64_t* a_empty = a_full + kTmaMaxStages;
  uint64_t* b_full = a_empty + kTmaMaxStages;
  uint64_t* b_empty = b_full + kTmaMaxStages;
  uint64_t* t_full = b_empty + kTmaMaxStages;
  uint64_t* t_empty = t_full + 4;
  uint64_t* b_res = t_empty + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_res + 1);
  volatile uint32_t* sk_last = tmem_slot + 2;          // split-K: "this CTA holds the last slice of the tile" (epilogue warps)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TMA_STAMP(0, threadIdx.x == 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kTmaMaxStages; ++i) {
      mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4 * C::kSlots); }
    mbar_init(b_res, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); }
  if (warp == 2) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  TMA_STAMP(1, threadIdx.x == 0);
  pdl_wait();            // everything above overlapped the previous kernel; its results are visible from here on
  pdl_trigger();
  TMA_STAMP(2, threadIdx.x == 0);

  const int tiles_per_img = g.tiles_x * g.tiles_y;
  const int total = g.n_img * tiles_per_img * g.n_blocks * (SK ? g.ksplit : 1);
  const int n_cs = g.chunks * T;

  if (warp == 0) {
    // ===================== TMA producer: the whole warp runs the (uniform) loops, one elected lane issues =========
    if (g.nbr == 0) {         // resident weights: every (chunk, s, r) tile once, slot = (chunk * T + s) * T + r
      if (elect_one()) {
        mbar_arrive_expect_tx(b_res, (uint32_t)(b_slots * C::kBStage));
        for (int slot = 0; slot < b_slots; ++slot) {
          const int r = slot % T, cs = slot / T, sft = cs % T, chunk = cs / T;
          tma_load_3d(b_tiles + (size_t)slot * C::kBStage, &map_b, b_res, (r * T + sft) * g.cin + chunk * KC, 0, 0);
        }
      }
      __syncwarp();
    }
    int sa = 0, pa = 0, sb = 0, pb = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      int tq = t, cs0 = 0, cs1 = n_cs;
      if constexpr (SK) {
        tq = g.div_ks.quot(t);
        cs0 = (t - tq * g.ksplit) * g.cs_per;
        cs1 = min(n_cs, cs0 + g.cs_per);
      }
      const int mt = g.div_nb.quot(tq), nb = tq - mt * g.n_blocks;
      const int img = g.div_tpi.quot(mt), rem = mt - img * tiles_per_img;
      const int ty = g.div_tx.quot(rem);
      const int oh0 = ty * kTmaTH + g.org, ow0 = (rem - ty * g.tiles_x) * kTmaTW + g.org;
      const int n0 = nb * BN;
      for (int chunk = SK ? cs0 / T : 0; chunk < (SK ? (cs1 + T - 1) / T : g.chunks); ++chunk) {
        for (int sft = 0; sft < T; ++sft) {
          if constexpr (SK) {
            if (chunk * T + sft < cs0 || chunk * T + sft >= cs1) continue;
          }
          if (!TMA_DBG(g, 8)) {
            mbar_wait(&a_empty[sa], (uint32_t)(pa ^ 1));
            if (elect_one()) {
              mbar_arrive_expect_tx(&a_full[sa], C::kAStage);
              tma_load_5d(a_tiles + sa * C::kAStage, &map_a, &a_full[sa], chunk * KC, ow0 + sft, oh0, img, 0);
            }
            __syncwarp();
            TMA_STAMP(3, lane == 0 && t == (int)blockIdx.x && chunk * T + sft == cs0);
            TMA_STAMP(4, lane == 0 && t == (int)blockIdx.x);
            if (++sa == g.na) { sa = 0; pa ^= 1; }
          }
          if (g.nbr > 0 && !TMA_DBG(g, 4)) {
            for (int r = 0; r < T; ++r) {
              mbar_wait(&b_empty[sb], (uint32_t)(pb ^ 1));
              if (elect_one()) {
                mbar_arrive_expect_tx(&b_full[sb], C::kBStage);
                tma_load_3d(b_tiles + sb * C::kBStage, &map_b, &b_full[sb], (r * T + sft) * g.cin + chunk * KC, n0, 0);
              }
              __syncwarp();
              if (++sb == g.nbr) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform loops (descriptors stay in uniform registers), one elected lane issues
    constexpr uint32_t idesc = umma_idesc_bf16(BN);
    constexpr uint32_t idesc2 = umma_idesc_bf16(2 * BN);      // N-concatenated weight tile
    constexpr uint32_t kLbo = 1u << 16;
    const uint32_t a_ring = ((smem_u32(a_tiles) >> 4) & 0x3FFF) | kLbo;       // low descriptor words: start address >> 4
    const uint32_t b_ring = ((smem_u32(b_tiles) >> 4) & 0x3FFF) | kLbo;
    const bool resident = g.nbr == 0;
    if (resident) { mbar_wait(b_res, 0); tc_fence_after(); }
    int sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int as = it % C::kAcc;
      mbar_wait(&t_empty[as], (uint32_t)(((it / C::kAcc) & 1) ^ 1));      // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * C::kAccCols);
      uint32_t accumulate = 0;
      int cs_first = 0, cs_last = n_cs;
      if constexpr (SK) {
        cs_first = (t - g.div_ks.quot(t) * g.ksplit) * g.cs_per;
        cs_last = min(n_cs, cs_first + g.cs_per);
      }
      for (int cs = cs_first; cs < (TMA_DBG(g, 512) ? 0 : cs_last); ++cs) {
        if (!TMA_DBG(g, 8)) { mbar_wait(&a_full[sa], (uint32_t)pa); tc_fence_after(); }
        TMA_STAMP(5, lane == 0 && it == 0 && cs == cs_first);
        const uint32_t a_st = a_ring + (uint32_t)(sa * (C::kAStage >> 4));
        if (resident) {
          // weights resident: nothing to wait for inside the stage, one election issues all T taps
          if (elect_one()) {
#pragma unroll
            for (int r = 0; r < T; ++r) {
              const uint32_t b_st = b_ring + (uint32_t)((cs * T + r) * (C::kBStage >> 4));
              const uint32_t a_hi = a_st + (uint32_t)((r * kTmaTW * C::kRowBytes) >> 4);
#pragma unroll
              for (int j = 0; j < KC / 16; ++j) {
                const uint32_t o = (uint32_t)(j * 2);
                const uint64_t dah = make_desc(C::kDescHi, a_hi + o), dal = make_desc(C::kDescHi, a_hi + o + (C::kAPlane >> 4));
                const uint64_t dbh = make_desc(C::kDescHi, b_st + o), dbl = make_desc(C::kDescHi, b_st + o + (C::kBPlane >> 4));
                if (!TMA_DBG(g, 1)) {
                  if constexpr (C::kNcat) {
                    umma_bf16(tmem_d, dah, dbh, idesc2, (r == 0 && j == 0) ? accumulate : 1u);     // A_hi.[B_hi; B_lo]
                    umma_bf16(tmem_d, dal, dbh, idesc, 1u);                                         // + A_lo.B_hi
                  } else {
                    umma_bf16(tmem_d, dal, dbh, idesc, (r == 0 && j == 0) ? accumulate : 1u);
                    umma_bf16(tmem_d, dah, dbl, idesc, 1u);
                    umma_bf16(tmem_d, dah, dbh, idesc, 1u);
                  }
                }
              }
            }
            umma_commit(&a_empty[sa]);
          }
          __syncwarp();
          accumulate = 1u;
          if (++sa == g.na) { sa = 0; pa ^= 1; }
          continue;
        }
#pragma unroll
        for (int r = 0; r < T; ++r) {
          mbar_wait(&b_full[sb], (uint32_t)pb);
          tc_fence_after();
          TMA_STAMP(6, lane == 0 && it == 0 && cs == cs_first && r == 0);
          const uint32_t b_st = b_ring + (uint32_t)(sb * (C::kBStage >> 4));
          const uint32_t a_hi = a_st + (uint32_t)((r * kTmaTW * C::kRowBytes) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < KC / 16; ++j) {
              const uint32_t o = (uint32_t)(j * 2);                    // 16 bf16 = 32 bytes along k inside the swizzled row
              const uint64_t dah = make_desc(C::kDescHi, a_hi + o), dal = make_desc(C::kDescHi, a_hi + o + (C::kAPlane >> 4));
              const uint64_t dbh = make_desc(C::kDescHi, b_st + o), dbl = make_desc(C::kDescHi, b_st + o + (C::kBPlane >> 4));
              if (!TMA_DBG(g, 1)) {
                if constexpr (C::kNcat) {
                  umma_bf16(tmem_d, dah, dbh, idesc2, (j == 0) ? accumulate : 1u);
                  umma_bf16(tmem_d, dal, dbh, idesc, 1u);
                } else {
                  umma_bf16(tmem_d, dal, dbh, idesc, (j == 0) ? accumulate : 1u);
                  umma_bf16(tmem_d, dah, dbl, idesc, 1u);
                  umma_bf16(tmem_d, dah, dbh, idesc, 1u);
                }
              }
            }
            umma_commit(&b_empty[sb]);
          }
          __syncwarp();
          accumulate = 1u;
          if (++sb == g.nbr) { sb = 0; pb ^= 1; }
        }
        if (!TMA_DBG(g, 8)) {
          if (elect_one()) umma_commit(&a_empty[sa]);
          __syncwarp();
          if (++sa == g.na) { sa = 0; pa ^= 1; }
        }
      }
      if (elect_one()) umma_commit(&t_full[as]);
      __syncwarp();
      TMA_STAMP(7, lane == 0 && it == 0);
    }
  } else {
    // ===================== epilogue =====================
    // 4 TMEM lane quarters x up to 4 column slots of 8 accumulator columns; with BN < 32 the surplus warps have no columns
    // and stay out of the hand-off (t_empty counts the working warps only)
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int slot = (warp - 2) >> 2;            // column slot: groups slot, slot + 4, ... of 8 accumulator columns
    const int l = q * 32 + lane;
    if (slot < C::kSlots) {
      // loop invariants of the plane output (bf16 elements)
      const long long o_px = g.out.c_total, o_row = (long long)g.out.wp * g.out.c_total;
      const long long o_img = (long long)g.out.hp * o_row;
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int as = it % C::kAcc;
        int tq = t;
        float* part = nullptr;
        if constexpr (SK) {
          tq = g.div_ks.quot(t);
          part = g.part + (long long)(t - tq * g.ksplit) * g.part_stride;
        }
        const int mt = g.div_nb.quot(tq), nb = tq - mt * g.n_blocks;
        const int img0 = g.div_tpi.quot(mt), rem = mt - img0 * tiles_per_img;
        const int ty = g.div_tx.quot(rem);
        const int oh0 = ty * kTmaTH + (l >> 3), ow = (rem - ty * g.tiles_x) * kTmaTW + (l & 7);
        // TALL: the tile walks the images stacked with their halo rows (see TmaConvArgs::tall_hp); row -> (image, row)
        const int img = TALL ? g.div_hp.quot(oh0) : img0;
        const int oh = TALL ? oh0 - img * g.tall_hp : oh0;
        const bool ok = oh < g.ho && ow < g.wo && (!TALL || img < g.tall_n);
        const size_t y_off = (((size_t)img * g.ho + oh) * g.wo + ow) * g.cout;
        // planes: primary position(s) by pointer arithmetic; the reflection-halo copies of border pixels on a slow path
        const int up = g.out.up;
        __nv_bfloat16* o_base = g.out.base + img * o_img + (long long)(up * oh + 1) * o_row + (long long)(up * ow + 1) * o_px;
        const bool edge = g.out.base != nullptr && g.out.halo != 0 && ok &&
                          (up == 1 ? (oh == 1 || oh == g.out.h_lo - 2 || ow == 1 || ow == g.out.w_lo - 2)
                                   : (oh == 0 || oh == g.out.h_lo - 1 || ow == 0 || ow == g.out.w_lo - 1));
        if (TMA_DBG(g, 256)) {
          mbar_wait(&t_full[as], (uint32_t)((it / C::kAcc) & 1));
        } else {
          if (lane == 0) mbar_wait(&t_full[as], (uint32_t)((it / C::kAcc) & 1));      // one poller per warp
          __syncwarp();
        }
        if (!TMA_DBG(g, 128)) tc_fence_after();
        TMA_STAMP(8, warp == 2 && lane == 0 && it == 0);
        if constexpr (SK) {
          if (g.cluster_ks > 0) {
            // cluster split-K: the raw accumulator tile of this K slice goes to this CTA's shared memory (the operand rings are
            // idle: the accumulator is complete, and a clustered launch gives every CTA exactly one work item)
            float* stage = reinterpret_cast<float*>(a_tiles) + (size_t)l * (BN + 4);
#pragma unroll 1
            for (int c = slot * 8; c < BN; c += 32) {
              uint32_t r[8];
              tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * C::kAccCols + c), r);
              if constexpr (C::kNcat) {
                uint32_t r2[8];
                tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * C::kAccCols + BN + c), r2);
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
              }
              *reinterpret_cast<uint4*>(stage + c) = make_uint4(r[0], r[1], r[2], r[3]);
              *reinterpret_cast<uint4*>(stage + c + 4) = make_uint4(r[4], r[5], r[6], r[7]);
            }
            tc_fence_before();
            continue;
          }
        }
        // scale / shift / residual / activation / gate, then fp32 and / or planes (primary cell, nearest-2x footprint, halo)
        auto finish = [&](float (&v)[8], int n) {
          if (g.scale) {
            const float4 s0 = ld4(g.scale + n), s1 = ld4(g.scale + n + 4);
            v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w; v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
          }
          if (g.shift) {
            const float4 s0 = ld4(g.shift + n), s1 = ld4(g.shift + n + 4);
            v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w; v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
          }
          if (g.res.base) {
            const __nv_bfloat16* rp = g.res.base + (((size_t)img * g.res.hp + oh + 1) * g.res.wp + ow + 1) * g.res.c_total + n;
            const uint4 h = *reinterpret_cast<const uint4*>(rp);
            const uint4 lo4 = *reinterpret_cast<const uint4*>(rp + g.res.plane);
            const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {lo4.x, lo4.y, lo4.z, lo4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = join_bf16x2(hh[j], ll[j]);
              v[2 * j] += f.x; v[2 * j + 1] += f.y;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = act_fast(v[j], g.act);
          if (g.gate) {
            const float4 g0 = ld4(g.gate + y_off + n), g1 = ld4(g.gate + y_off + n + 4);
            v[0] *= act_grad_from_output(g0.x, g.gate_act); v[1] *= act_grad_from_output(g0.y, g.gate_act);
            v[2] *= act_grad_from_output(g0.z, g.gate_act); v[3] *= act_grad_from_output(g0.w, g.gate_act);
            v[4] *= act_grad_from_output(g1.x, g.gate_act); v[5] *= act_grad_from_output(g1.y, g.gate_act);
            v[6] *= act_grad_from_output(g1.z, g.gate_act); v[7] *= act_grad_from_output(g1.w, g.gate_act);
          }
          if (g.y) {
            float* p = g.y + y_off + n;
            *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
          if (g.out.base) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
            const uint4 hv = make_uint4(hi[0], hi[1], hi[2], hi[3]), lv = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            __nv_bfloat16* p = o_base + n;
            *reinterpret_cast<uint4*>(p) = hv;
            *reinterpret_cast<uint4*>(p + g.out.plane) = lv;
            if (up == 2) {                        // nearest 2x: the other three pixels of the footprint
              *reinterpret_cast<uint4*>(p + o_px) = hv;             *reinterpret_cast<uint4*>(p + o_px + g.out.plane) = lv;
              *reinterpret_cast<uint4*>(p + o_row) = hv;            *reinterpret_cast<uint4*>(p + o_row + g.out.plane) = lv;
              *reinterpret_cast<uint4*>(p + o_row + o_px) = hv;     *reinterpret_cast<uint4*>(p + o_row + o_px + g.out.plane) = lv;
            }
            if (edge) {                           // border pixels also feed the halo cells that mirror onto them
              int rows[3], cols[3];
              plane_coords(oh, g.out.h_lo, up, 1, rows);
              plane_coords(ow, g.out.w_lo, up, 1, cols);
#pragma unroll 1
              for (int a = 0; a < 3; ++a) {
#pragma unroll 1
                for (int b = 0; b < 3; ++b) {
                  const int ra = a == 0 ? rows[0] : (a == 1 ? rows[1] : rows[2]);
                  const int cb = b == 0 ? cols[0] : (b == 1 ? cols[1] : cols[2]);
                  const bool primary = up == 1 ? (a == 0 && b == 0) : (a < 2 && b < 2);
                  if (ra < 0 || cb < 0 || primary) continue;
                  __nv_bfloat16* ph = g.out.base + img * o_img + ra * o_row + cb * o_px + n;
                  *reinterpret_cast<uint4*>(ph) = hv;
                  *reinterpret_cast<uint4*>(ph + g.out.plane) = lv;
                }
              }
            }
          }
        };
#pragma unroll 1
        for (int c = slot * 8; c < BN; c += 32) {
          uint32_t r[8];
          tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * C::kAccCols + c), r);
          TMA_STAMP(9, warp == 2 && lane == 0 && it == 0 && c == slot * 8);
          if constexpr (C::kNcat) {               // + the A_hi.B_lo block
            uint32_t r2[8];
            tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * C::kAccCols + BN + c), r2);
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
          }
          const int n = nb * BN + c;
          if (!ok || n >= g.cout || TMA_DBG(g, 2)) continue;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
          if constexpr (SK) {                     // raw partial sums of this K slice
            float* p = part + y_off + n;
            *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
            continue;
          }
          finish(v, n);
        }
        if constexpr (SK) {
          if (g.counters != nullptr) {
            // the accumulator has been read: hand it back before the (possibly long) reduction below
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[as]);
            // publish this CTA's slice, then find out whether it was the last one of the tile
            __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kTmaEpiWarps) : "memory");
            if (warp == 2 && lane == 0) {
              const int old = atomicAdd(&g.counters[tq], 1);
              *sk_last = (old == g.ksplit - 1) ? 1u : 0u;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kTmaEpiWarps) : "memory");
            if (*sk_last != 0u) {
              __threadfence();                    // the other slices are visible (their CTAs fenced before incrementing)
#pragma unroll 1
              for (int c = slot * 8; c < BN; c += 32) {
                const int n = nb * BN + c;
                if (!ok || n >= g.cout) continue;
                const float* p0 = g.part + y_off + n;
                float v[8];
                {
                  const float4 a = __ldcg(reinterpret_cast<const float4*>(p0)), b = __ldcg(reinterpret_cast<const float4*>(p0 + 4));
                  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                }
                for (int z = 1; z < g.ksplit; ++z) {          // fixed order: deterministic whichever CTA reduces
                  const float* pz = p0 + (size_t)z * g.part_stride;
                  const float4 a = __ldcg(reinterpret_cast<const float4*>(pz)), b = __ldcg(reinterpret_cast<const float4*>(pz + 4));
                  v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
                }
                finish(v, n);
              }
              if (warp == 2 && lane == 0) g.counters[tq] = 0;          // ready for the next launch / graph replay
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kTmaEpiWarps) : "memory");          // sk_last is rewritten by the next item
            continue;
          }
        }
        if (!TMA_DBG(g, 128)) tc_fence_before();
        __syncwarp();
        TMA_STAMP(10, warp == 2 && lane == 0 && it == 0);
        if (lane == 0) mbar_arrive(&t_empty[as]);
      }
    }
  }
  if constexpr (SK) {
    if (g.cluster_ks > 0) {
      // every thread of every CTA of the cluster: the staged tiles are complete and visible cluster-wide
      cluster_sync_all();
      if (warp >= 2) {
        const int cks = g.cluster_ks;
        const uint32_t rank = cluster_ctarank();
        const int tq = (int)blockIdx.x / cks;
        const int mt = g.div_nb.quot(tq), nb = tq - mt * g.n_blocks;
        const int img0 = g.div_tpi.quot(mt), rem = mt - img0 * tiles_per_img;
        const int ty = g.div_tx.quot(rem);
        const int rows = kTcBM / cks;                        // accumulator rows (= output pixels) this CTA finishes
        constexpr int kGroups = BN / 8;
        const uint32_t stage0 = smem_u32(a_tiles);
        for (int u = (int)threadIdx.x - 64; u < rows * kGroups; u += 32 * kTmaEpiWarps) {
          const int px = (int)rank * rows + u / kGroups, c = (u % kGroups) * 8;
          const int oh0 = ty * kTmaTH + (px >> 3), ow = (rem - ty * g.tiles_x) * kTmaTW + (px & 7);
          const int img = TALL ? g.div_hp.quot(oh0) : img0;
          const int oh = TALL ? oh0 - img * g.tall_hp : oh0;
          const int n = nb * BN + c;
          if (!(oh < g.ho && ow < g.wo && (!TALL || img < g.tall_n)) || n >= g.cout) continue;
          const uint32_t addr = stage0 + (uint32_t)((px * (BN + 4) + c) * 4);
          float v[8];
          {
            const float4 a = ld_dsmem4(addr, 0), b = ld_dsmem4(addr + 16, 0);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
          }
          for (int z = 1; z < cks; ++z) {                    // fixed rank order: the sum does not depend on timing
            const float4 a = ld_dsmem4(addr, (uint32_t)z), b = ld_dsmem4(addr + 16, (uint32_t)z);
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
          }
          tma_epilogue_store(g, img, oh, ow, n, v);
        }
      }
      cluster_sync_all();          // nobody leaves while a peer may still read its shared memory
    }
  }
  TMA_STAMP(11, threadIdx.x == 0);
  tc_fence_before();
  __syncthreads();
  TMA_STAMP(12, threadIdx.x == 0);
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
  TMA_STAMP(13, warp == 2 && lane == 0);
}

// ---- split-K epilogue --------------------------------------------------------------------------------
// y = epilogue(sum_z part[z]) for the split-K form of conv3x3_tma_kernel: the z slices are added in a fixed order
// (deterministic), then the same scale / shift / residual / activation / gate sequence as the fused epilogue above;
// fp32 NHWC and / or planes (reflection halo, nearest 2x on write).  One thread = 8 channels of one output pixel.
struct TmaSplitkEpiArgs {
  int n_img, ho, wo, cout, ksplit;
  long long part_stride;
  const float* part;
  int act;
  const float* scale;
  const float* shift;
  PlaneView res;
  const float* gate;
  int gate_act;
  float* y;
  PlaneView out;
};

__global__ void __launch_bounds__(256)
tma_splitk_epilogue_kernel(const TmaSplitkEpiArgs g) {
  pdl_wait();
  pdl_trigger();
  const int c8 = g.cout / 8;
  const size_t total = (size_t)g.n_img * g.ho * g.wo * c8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % c8) * 8;
    size_t px = i / c8;
    const int ow = (int)(px % g.wo); px /= g.wo;
    const int oh = (int)(px % g.ho);
    const int img = (int)(px / g.ho);
    const size_t off = i * 8;                       // = ((img * ho + oh) * wo + ow) * cout + n
    float v[8];
    {
      const float4 a = ld4(g.part + off), b = ld4(g.part + off + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    for (int z = 1; z < g.ksplit; ++z) {
      const float* pz = g.part + (size_t)z * g.part_stride + off;
      const float4 a = ld4(pz), b = ld4(pz + 4);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (g.scale) {
      const float4 s0 = ld4(g.scale + n), s1 = ld4(g.scale + n + 4);
      v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w; v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
    }
    if (g.shift) {
      const float4 s0 = ld4(g.shift + n), s1 = ld4(g.shift + n + 4);
      v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w; v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
    }
    if (g.res.base) {
      float r[8];
      planes_load8(g.res, img, oh + 1, ow + 1, n, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = act_fast(v[j], g.act);
    if (g.gate) {
      const float4 g0 = ld4(g.gate + off), g1 = ld4(g.gate + off + 4);
      v[0] *= act_grad_from_output(g0.x, g.gate_act); v[1] *= act_grad_from_output(g0.y, g.gate_act);
      v[2] *= act_grad_from_output(g0.z, g.gate_act); v[3] *= act_grad_from_output(g0.w, g.gate_act);
      v[4] *= act_grad_from_output(g1.x, g.gate_act); v[5] *= act_grad_from_output(g1.y, g.gate_act);
      v[6] *= act_grad_from_output(g1.z, g.gate_act); v[7] *= act_grad_from_output(g1.w, g.gate_act);
    }
    if (g.y) {
      *reinterpret_cast<float4*>(g.y + off) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(g.y + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (g.out.base) planes_emit<8>(g.out, img, oh, ow, n, v);
  }
}

// ---- fp32 NHWC -> planes -------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
planes_from_f32_kernel(const float* __restrict__ x, PlaneView out, int n, int h, int w, int c) {
  pdl_wait();
  pdl_trigger();
  const int c4 = c / 4;
  const size_t total = (size_t)n * h * w * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c4) * 4;
    size_t p = i / c4;
    const int ow = (int)(p % w); p /= w;
    const int oh = (int)(p % h);
    const int img = (int)(p / h);
    const float4 v4 = ld4(x + i * 4);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    planes_emit<4>(out, img, oh, ow, cc, v);
  }
}

// ---- conv1 in space-to-depth form --------------------------------------------------------------------
// (x - 0.45) / 0.225 (networks/resnet_encoder.py:117) of one or two NCHW images, 2x2 space-to-depth, channel padding
// 3 -> 4 per image, written as planes: channel (p * 2 + q) * cp + c of S2D pixel (y, x) = image pixel (2y + p, 2x + q).
// A 7x7 / stride-2 / pad-3 convolution of the image is the 4x4 (rows o-2 .. o+1) stride-1 convolution of this tensor.
__global__ void __launch_bounds__(256)
stage_s2d_kernel(const float* __restrict__ a, const float* __restrict__ b, PlaneView out, int n, int h, int w, int img0) {
  pdl_wait();
  pdl_trigger();
  const int h2 = h / 2, w2 = w / 2, cp = b ? 8 : 4;
  const size_t total = (size_t)n * h2 * w2 * 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int pq = (int)(i & 3);
    size_t t = i >> 2;
    const int x2 = (int)(t % w2); t /= w2;
    const int y2 = (int)(t % h2);
    const int img = (int)(t / h2);
    const size_t hw = (size_t)h * w;
    const size_t px = (size_t)(2 * y2 + (pq >> 1)) * w + 2 * x2 + (pq & 1);
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = 0.f;
    const float* pa = a + (size_t)img * 3 * hw + px;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (ld1(pa + c * hw) - 0.45f) / 0.225f;
    if (b) {
      const float* pb = b + (size_t)img * 3 * hw + px;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[3 + c] = (ld1(pb + c * hw) - 0.45f) / 0.225f;
      planes_emit<8>(out, img0 + img, y2, x2, pq * cp, v);
    } else {
      const float v4[4] = {v[0], v[1], v[2], v[3]};
      planes_emit<4>(out, img0 + img, y2, x2, pq * cp, v4);
    }
  }
}

// fp32 NHWC (n, h, w, c) -> 2x2 space-to-depth planes (n, h/2, w/2, 4c): channel (p*2 + q) * c + k of S2D pixel (y, x) is
// channel k of pixel (2y + p, 2x + q).  A 3x3 / stride-2 / pad-1 convolution of the tensor is the 2x2 (rows o-1 .. o)
// stride-1 convolution of this one with W'[co][r][s][(p, q, k)] = W[co][2r + p - 1][2s + q - 1][k] (zero for index -1).
__global__ void __launch_bounds__(256)
s2d_planes_kernel(const float* __restrict__ x, PlaneView out, int n, int h, int w, int c) {
  pdl_wait();
  pdl_trigger();
  const int c4 = c / 4;
  const size_t total = (size_t)n * h * w * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c4) * 4;
    size_t p = i / c4;
    const int xw = (int)(p % w); p /= w;
    const int yh = (int)(p % h);
    const int img = (int)(p / h);
    const float4 v4 = ld4(x + i * 4);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    planes_emit<4>(out, img, yh >> 1, xw >> 1, ((yh & 1) * 2 + (xw & 1)) * c + cc, v);
  }
}

// ---- dgrad helpers ---------------------------------------------------------------------------------
// Wt[ci'][2 - r][2 - s][co] = W[co][r][s][ci'] for ci' < n_ci, on both bf16 planes (a pure permutation): the weights
// of the convolution that computes dX from dZ (taps flipped, channels transposed)
struct WtFlipItem { long long src, dst; int co, ci, n_ci, pad; };

__global__ void __launch_bounds__(256)
wt_flip_kernel(const WtFlipItem* __restrict__ items, const uint16_t* __restrict__ w_hi, long long w_plane,
               uint16_t* __restrict__ wt_hi, long long wt_plane) {
  pdl_wait();
  pdl_trigger();
  const WtFlipItem it = items[blockIdx.y];
  const long long total = (long long)it.n_ci * 9 * it.co;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % it.co);
    const int tap = (int)((i / it.co) % 9);
    const int c = (int)(i / ((long long)it.co * 9));
    const long long src = it.src + ((long long)o * 9 + (8 - tap)) * it.ci + c;
    wt_hi[it.dst + i] = w_hi[src];
    wt_hi[wt_plane + it.dst + i] = w_hi[w_plane + src];
  }
}

// dX from the "full" correlation dxp (n, H + 2, W + 2, c) of a reflection-padded layer: every interior pixel also
// receives the halo cells that mirror onto it; up = 2 additionally sums the 2 x 2 nearest-neighbour footprint (the
// layer read an up-sampled tensor).  Then + the dispconv input gradient, * act'(y_prev); written as fp32 and/or planes.
// SK: dxp is the split-K workspace of the preceding conv3x3_tma_kernel<.., true> ([ksplit][n][H + 2][W + 2][c] raw partial
// sums, slices `part_stride` elements apart): the fold also adds the slices (fixed order), replacing the split-K epilogue
template <bool SK>
__device__ __forceinline__ void dgrad_fold_body(const float* __restrict__ dxp, int n, int H, int W, int c, int up,
                                                const float* __restrict__ y_prev, int act_prev, const float* __restrict__ disp_dz,
                                                const float* __restrict__ disp_w, float* __restrict__ dz_prev, PlaneView po,
                                                int ksplit, long long part_stride) {
  const int ho = H / up, wo = W / up, c4 = c / 4;
  const size_t total = (size_t)n * ho * wo * c4;
  const int Wp = W + 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c4) * 4;
    size_t p = i / c4;
    const int ow = (int)(p % wo); p /= wo;
    const int oh = (int)(p % ho);
    const int img = (int)(p / ho);
    float4 acc = f4zero();
    const int h0 = oh * up, w0 = ow * up;
    if (h0 >= 2 && h0 + up <= H - 2 && w0 >= 2 && w0 + up <= W - 2) {
      // interior: no halo cell mirrors onto this pixel -- the up x up cells, all loads issued before the first add
      const size_t base = (((size_t)img * (H + 2) + h0 + 1) * Wp + w0 + 1) * c + cc;
      if constexpr (SK) {
        for (int z = 0; z < ksplit; ++z) {
          const float* pz = dxp + (size_t)z * part_stride + base;
          float4 v0 = ld4(pz), v1 = f4zero(), v2 = f4zero(), v3 = f4zero();
          if (up == 2) { v1 = ld4(pz + c); v2 = ld4(pz + (size_t)Wp * c); v3 = ld4(pz + (size_t)Wp * c + c); }
          f4add(v0, v1); f4add(v2, v3); f4add(v0, v2); f4add(acc, v0);
        }
      } else {
        float4 v0 = ld4(dxp + base), v1 = f4zero(), v2 = f4zero(), v3 = f4zero();
        if (up == 2) { v1 = ld4(dxp + base + c); v2 = ld4(dxp + base + (size_t)Wp * c); v3 = ld4(dxp + base + (size_t)Wp * c + c); }
        f4add(v0, v1); f4add(v2, v3); f4add(v0, v2);
        acc = v0;
      }
    } else
    for (int a = 0; a < up; ++a) {
      const int hh = oh * up + a;                       // interior row of the layer input
      int rows[3] = {hh + 1, hh == 1 ? 0 : -1, hh == H - 2 ? H + 1 : -1};
      for (int b = 0; b < up; ++b) {
        const int ww = ow * up + b;
        int cols[3] = {ww + 1, ww == 1 ? 0 : -1, ww == W - 2 ? W + 1 : -1};
#pragma unroll
        for (int ra = 0; ra < 3; ++ra) {
          if (rows[ra] < 0) continue;
#pragma unroll
          for (int cb = 0; cb < 3; ++cb) {
            if (cols[cb] < 0) continue;
            const size_t cell = (((size_t)img * (H + 2) + rows[ra]) * Wp + cols[cb]) * c + cc;
            if constexpr (SK) {
              for (int z = 0; z < ksplit; ++z) f4add(acc, ld4(dxp + (size_t)z * part_stride + cell));
            } else {
              f4add(acc, ld4(dxp + cell));
            }
          }
        }
      }
    }
    if (disp_dz) f4add(acc, dispconv_dgrad_at(disp_dz, disp_w, img, oh, ow, ho, wo, c, cc));
    if (y_prev) {
      const float4 yv = ld4(y_prev + i * 4);
      acc.x *= act_grad_from_output(yv.x, act_prev); acc.y *= act_grad_from_output(yv.y, act_prev);
      acc.z *= act_grad_from_output(yv.z, act_prev); acc.w *= act_grad_from_output(yv.w, act_prev);
    }
    if (dz_prev) *reinterpret_cast<float4*>(dz_prev + i * 4) = acc;
    if (po.base) {
      const float v[4] = {acc.x, acc.y, acc.z, acc.w};
      planes_emit<4>(po, img, oh, ow, cc, v);
    }
  }
}

__global__ void __launch_bounds__(256)
dgrad_fold_kernel(const float* __restrict__ dxp, int n, int H, int W, int c, int up, const float* __restrict__ y_prev,
                  int act_prev, const float* __restrict__ disp_dz, const float* __restrict__ disp_w,
                  float* __restrict__ dz_prev, PlaneView po) {
  pdl_wait();
  pdl_trigger();
  dgrad_fold_body<false>(dxp, n, H, W, c, up, y_prev, act_prev, disp_dz, disp_w, dz_prev, po, 1, 0);
}

__global__ void __launch_bounds__(256)
dgrad_fold_splitk_kernel(const float* __restrict__ part, int ksplit, long long part_stride, int n, int H, int W, int c, int up,
                         const float* __restrict__ y_prev, int act_prev, const float* __restrict__ disp_dz,
                         const float* __restrict__ disp_w, float* __restrict__ dz_prev, PlaneView po) {
  pdl_wait();
  pdl_trigger();
  dgrad_fold_body<true>(part, n, H, W, c, up, y_prev, act_prev, disp_dz, disp_w, dz_prev, po, ksplit, part_stride);
}

// ---- host side ---------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled g_encode = nullptr;

static int get_encoder() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    set_error("cuTensorMapEncodeTiled is not available (%s)", cudaGetErrorString(e));
    return -3;
  }
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  return 0;
}

static CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

static int encode(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                  int row_bytes, const char* what) {
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  const CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides, box, ones,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(%s) failed: %d", what, (int)r); return -4; }
  return 0;
}

static int g_tma_debug = 0;
static int g_tma_sms = 0;
static int tma_num_sms() {
  if (g_tma_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_tma_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_tma_sms <= 0)
      g_tma_sms = 148;
  }
  return g_tma_sms;
}

template <class... KArgs, class... Args>
static void launch_k_cluster(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);      // errors surface through finish_launch
}

template <int KC, int BN, int T, bool SK = false, bool TALL = false>
static int launch_tma_conv(const CUtensorMap& ma, const CUtensorMap& mb, TmaConvArgs g, cudaStream_t st) {
  using C = TmaConvCfg<KC, BN, T>;
  auto kern = conv3x3_tma_kernel<KC, BN, T, SK, TALL>;
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemMax);
    if (e != cudaSuccess) { set_error("conv3x3_tma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -3; }
    attr_set = true;
  }
  // shared-memory plan: weights resident when the whole tensor fits next to >= 3 A stages, else a B ring
  const int budget = kTmaSmemMax - 1024 - kTmaBarBytes;
  const long long w_bytes = (long long)g.chunks * T * T * C::kBStage;
  long long b_bytes;
  if (!SK && g.n_blocks == 1 && w_bytes <= kTmaResidentMax && budget - w_bytes >= 3LL * C::kAStage) {
    g.nbr = 0; b_bytes = w_bytes;
  } else {
    g.nbr = (KC == 64 && BN == 128) ? 3 : 6; b_bytes = (long long)g.nbr * C::kBStage;
  }
  long long na = (budget - b_bytes) / C::kAStage;
  g.na = (int)(na > kTmaMaxStages ? kTmaMaxStages : na);
  if (g.na < 2) { set_error("conv3x3_tma: shared memory plan failed"); return -3; }
  const int smem = (int)(g.na * C::kAStage + b_bytes + 1024 + kTmaBarBytes);
  const long long total = (long long)g.n_img * g.tiles_x * g.tiles_y * g.n_blocks * (SK ? g.ksplit : 1);
  const int cap = (g.grid_cap > 0 && g.grid_cap < tma_num_sms()) ? g.grid_cap : tma_num_sms();
  const int grid = (int)(total < cap ? total : cap);
  if constexpr (SK) {
    if (g.cluster_ks > 0) {          // one work item per CTA, the ksplit slices of a tile = one cluster
      if (total > tma_num_sms() || g.cluster_ks != g.ksplit || g.na * C::kAStage < kTcBM * (BN + 4) * 4) {
        set_error("conv3x3_tma: cluster split-K plan does not fit (%lld items, ksplit %d)", total, g.ksplit); return -3;
      }
      launch_k_cluster(kern, (int)total, kTmaThreads, smem, st, g.cluster_ks, ma, mb, g);
      return finish_launch("conv3x3_tma (cluster split-K)");
    }
  }
  launch_k(kern, dim3(grid), dim3(kTmaThreads), smem, st, ma, mb, g);
  return finish_launch(SK ? "conv3x3_tma (split-K)" : "conv3x3_tma");
}

// ---- split-K plan ------------------------------------------------------------------------------------
// Layers whose output grid has fewer (tile, N block) work items than half the SMs (everything <= 24 x 80 at batch 1)
// cannot fill the persistent kernel; their K = 9 * cin is long (cin 128..512), so the (channel chunk, column shift)
// steps are spread over ksplit work items per tile.  Built for the shapes those layers have: 64-channel chunks,
// N block 64 / 128, 3x3 and the 2x2 space-to-depth form.  CLSLAM_TMA_SPLITK_ITEMS overrides the item threshold.
static int tma_splitk_items() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CLSLAM_TMA_SPLITK_ITEMS");
    v = e ? atoi(e) : tma_num_sms() / 2;
    if (v < 0) v = 0;
  }
  return v;
}
// CLSLAM_TMA_SPLITK_WAVES=1 additionally splits launches with SMs / 2 < items < SMs (layer3 / layer4 at batch 8: 80 / 96
// items) when that shortens the schedule: cost = waves x steps per item (+ 2 steps for the extra epilogue launch),
// e.g. 80 items x 12 steps: 1 wave x 12 -> 3 slices: 2 waves x 4
static bool tma_splitk_waves() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CLSLAM_TMA_SPLITK_WAVES"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
struct TmaSplit { int ksplit, cs_per; };
static TmaSplit tma_split_plan(long long items, int n_cs, int kc, int bn, int taps, int budget = 0) {
  TmaSplit sp{1, n_cs};
  const bool built = kc == 64 && ((taps == 3 && (bn == 64 || bn == 128)) || (taps == 2 && bn == 128));
  if (!built || items < 1 || n_cs < 2) return sp;
  const int sms = (budget > 0 && budget < tma_num_sms()) ? budget : tma_num_sms();
  if (items > tma_splitk_items()) {
    if (!tma_splitk_waves() || items >= sms) return sp;
    long long best = n_cs;
    for (int ks = 2; ks <= 4 && ks <= n_cs; ++ks) {
      const int per = (n_cs + ks - 1) / ks, ks2 = (n_cs + per - 1) / per;
      if (ks2 < 2 || 5 * n_cs < 4 * ks2 * per) continue;
      const long long waves = (items * ks2 + sms - 1) / sms, cost = waves * per + 2;
      if (cost < best) { best = cost; sp.ksplit = ks2; sp.cs_per = per; }
    }
    return sp;
  }
  long long want = sms / items;                      // one wave of work items (within the launch's SM budget)
  if (want > n_cs) want = n_cs;
  if (want < 2) return sp;
  sp.cs_per = (int)((n_cs + want - 1) / want);
  sp.ksplit = (n_cs + sp.cs_per - 1) / sp.cs_per;     // every z slice has at least one step
  // not worth it when the slices are badly balanced (3 steps as 2 + 1: the 48 x 160 / 64-channel layers at batch 1)
  if (sp.ksplit < 2 || 5 * n_cs < 4 * sp.ksplit * sp.cs_per) { sp.ksplit = 1; sp.cs_per = n_cs; }
  return sp;
}

static int check_planes(const clslam_planes* p, const char* what) {
  if (!p || !p->base) { set_error("%s: null planes", what); return -1; }
  if (p->c_total % 8 || p->c_off % 8 || p->c % 8 || p->c_off + p->c > p->c_total || p->n < 1 || p->h < 2 || p->w < 2 ||
      (reinterpret_cast<uintptr_t>(p->base) & 15) || (p->plane_stride % 8)) {
    set_error("%s: bad planes (c_total=%d c_off=%d c=%d n=%d h=%d w=%d)", what, p->c_total, p->c_off, p->c, p->n, p->h, p->w);
    return -1;
  }
  return 0;
}

}  // namespace clslam

using namespace clslam;

// geometry shared by the launch and the workspace query
// Tall form (TmaConvArgs::tall_hp): > 0 = number of 16-row tiles when the images are walked as one stack, 0 = not applicable
// (the caller did not vouch for the halo, other kernel shapes, a single image) or no fewer tiles than one image at a time
static int tma_tall_tiles_y(const clslam_conv_tma* p, int taps, int kc, int bn) {
  if (!(p->flags & CLSLAM_TMA_IN_HALO_VALID) || (taps != 3 && taps != 2) || p->pad_mode == CLSLAM_PAD_FULL || p->in.n < 2) return 0;
  if (kc != 64 || bn != 128) return 0;                      // instantiated for the layers that have small maps
  const long long rows = (long long)p->in.n * (p->in.h + 2) - 2;         // output rows of the stack
  const long long tall = (rows + kTmaTH - 1) / kTmaTH;
  const long long plain = (long long)p->in.n * ((p->in.h + kTmaTH - 1) / kTmaTH);
  return tall < plain ? (int)tall : 0;
}

struct TmaGeom { int taps, kc, bn, ho, wo, tiles_x, tiles_y, n_blocks, chunks; long long items; };
static int tma_geometry(const clslam_conv_tma* p, TmaGeom& q) {
  if (!p) { set_error("conv3x3_tma: null"); return -1; }
  const int cin = p->in.c, cout = p->cout;
  if (cin < 16 || cin % 16 || cout < 16 || cout % 16) { set_error("conv3x3_tma: cin %% 16, cout %% 16"); return -1; }
  q.taps = p->taps == 0 ? 3 : p->taps;
  q.kc = cin % 64 == 0 ? 64 : (cin % 32 == 0 ? 32 : 16);
  q.bn = cout >= 128 ? 128 : (cout >= 64 ? 64 : (cout >= 32 ? 32 : 16));
  const int full = p->pad_mode == CLSLAM_PAD_FULL;
  q.ho = p->in.h + (full ? 2 : 0); q.wo = p->in.w + (full ? 2 : 0);
  q.tiles_x = (q.wo + kTmaTW - 1) / kTmaTW; q.tiles_y = (q.ho + kTmaTH - 1) / kTmaTH;
  q.n_blocks = cout / q.bn; q.chunks = cin / q.kc;
  q.items = (long long)p->in.n * q.tiles_x * q.tiles_y * q.n_blocks;
  if (const int ty = tma_tall_tiles_y(p, q.taps, q.kc, q.bn)) { q.tiles_y = ty; q.items = (long long)q.tiles_x * ty * q.n_blocks; }
  return 0;
}

// arguments of the fold that follows a PAD_FULL dgrad launch (clslam_dgrad_fold)
struct FoldArgs {
  int up;
  const float* y_prev; int act_prev;
  const float* disp_dz; const float* disp_w;
  float* dz_prev; const clslam_planes* dz_prev_planes;
};
static int launch_fold(const float* dxp, int ksplit, long long part_stride, int n, int h, int w, int c, const FoldArgs& f, cudaStream_t st);
static int conv3x3_tma_impl(const clslam_conv_tma* p, void* workspace, long long workspace_bytes, void* stream,
                            const FoldArgs* fold = nullptr, bool* folded = nullptr);

extern "C" {

int clslam_conv3x3_tma(const clslam_conv_tma* p, void* stream) { return conv3x3_tma_impl(p, nullptr, 0, stream); }

long long clslam_conv3x3_tma_workspace_bytes(const clslam_conv_tma* p) {
  TmaGeom q{};
  if (tma_geometry(p, q)) return -1;
  const TmaSplit sp = tma_split_plan(q.items, q.chunks * q.taps, q.kc, q.bn, q.taps, p->sm_budget);
  if (sp.ksplit < 2) return 0;
  return (long long)sp.ksplit * p->in.n * q.ho * q.wo * p->cout * 4;
}

int clslam_conv3x3_tma_ws(const clslam_conv_tma* p, void* workspace, long long workspace_bytes, void* stream) {
  return conv3x3_tma_impl(p, workspace, workspace_bytes, stream);
}

}  // extern "C"

static int conv3x3_tma_impl(const clslam_conv_tma* p, void* workspace, long long workspace_bytes, void* stream, const FoldArgs* fold,
                            bool* folded) {
  if (!p) { set_error("conv3x3_tma: null"); return -1; }
  if (int rc = check_planes(&p->in, "conv3x3_tma.in")) return rc;
  const clslam_planes& in = p->in;
  const int cin = in.c, cout = p->cout;
  if (cin % 16 || cout % 16 || !p->w_hi || !p->w_lo) { set_error("conv3x3_tma: cin %% 16, cout %% 16, weights"); return -1; }
  if (p->pad_mode < 0 || p->pad_mode > 2) { set_error("conv3x3_tma: pad_mode"); return -1; }
  const int taps = p->taps == 0 ? 3 : p->taps;
  if (taps != 1 && taps != 2 && taps != 3 && taps != 4) { set_error("conv3x3_tma: taps must be 1, 2, 3 or 4"); return -1; }
  if (taps != 3 && p->pad_mode != CLSLAM_PAD_ZERO) { set_error("conv3x3_tma: the 1x1 / 2x2 / 4x4 forms are zero padded"); return -1; }
  if (int rc = get_encoder()) return rc;
  const int kc = cin % 64 == 0 ? 64 : (cin % 32 == 0 ? 32 : 16);
  const int bn = cout >= 128 ? 128 : (cout >= 64 ? 64 : (cout >= 32 ? 32 : 16));
  if (cout % bn) { set_error("conv3x3_tma: cout %d is not a multiple of the N block %d", cout, bn); return -1; }
  const int full = p->pad_mode == CLSLAM_PAD_FULL;
  const int ho = in.h + (full ? 2 : 0), wo = in.w + (full ? 2 : 0);

  TmaConvArgs g{};
  g.n_img = in.n; g.ho = ho; g.wo = wo; g.cout = cout; g.cin = cin;
  g.tiles_x = (wo + kTmaTW - 1) / kTmaTW; g.tiles_y = (ho + kTmaTH - 1) / kTmaTH;
  g.n_blocks = cout / bn; g.chunks = cin / kc;
  g.div_nb = make_fastdiv(g.n_blocks); g.div_tpi = make_fastdiv(g.tiles_x * g.tiles_y); g.div_tx = make_fastdiv(g.tiles_x);
  g.debug = g_tma_debug;
  g.grid_cap = p->sm_budget;
  g.gate = p->gate; g.gate_act = p->gate_act;
  g.act = p->act; g.scale = p->scale; g.shift = p->shift; g.y = p->y;
  if (p->residual) {
    if (int rc = check_planes(p->residual, "conv3x3_tma.residual")) return rc;
    if (p->residual->h != ho || p->residual->w != wo || p->residual->c != cout || p->residual->n != in.n) { set_error("conv3x3_tma: residual dims"); return -1; }
    g.res = make_plane_view(p->residual, 1, 0);
  }
  if (p->out) {
    if (int rc = check_planes(p->out, "conv3x3_tma.out")) return rc;
    const int up = p->out_up;
    if ((up != 1 && up != 2) || p->out->h != ho * up || p->out->w != wo * up || p->out->c != cout || p->out->n != in.n) {
      set_error("conv3x3_tma: out dims (%d x %d x %d, up %d) vs (%d x %d x %d)", p->out->h, p->out->w, p->out->c, up, ho, wo, cout);
      return -1;
    }
    g.out = make_plane_view(p->out, up, p->out_halo);
  }

  // A: 5-D map (channel, x, y, image, plane) over the padded view (reflect: the stored halo is the padding) or the
  // interior view (zero / full: TMA zero-fills outside)
  const int hp = in.h + 2, wp = in.w + 2;
  // tall form: the n images are one stack of n * hp rows whose halo rows / columns are the padding (EXPERIMENTAL, opt-in
  // through p->flags; not yet run on hardware)
  const int tall_ty = tma_tall_tiles_y(p, taps, kc, bn);
  const bool tall = tall_ty > 0;
  if (tall) {
    g.n_img = 1; g.tiles_y = tall_ty;
    g.div_tpi = make_fastdiv(g.tiles_x * g.tiles_y);
    g.tall_hp = hp; g.tall_n = in.n; g.div_hp = make_fastdiv(hp);
  }
  const bool padded_view = p->pad_mode == CLSLAM_PAD_REFLECT || tall;
  __nv_bfloat16* abase = reinterpret_cast<__nv_bfloat16*>(in.base) + in.c_off + (padded_view ? 0 : ((size_t)wp + 1) * in.c_total);
  g.org = padded_view ? 0 : ((full || taps == 4) ? -2 : (taps == 1 ? 0 : -1));      // (tall 2x2: padded rows v, v + 1 = input rows v - 1, v)      // 4x4: input rows o - 2 .. o + 1; 2x2: o - 1 .. o; 1x1: o
  const cuuint64_t adims[5] = {(cuuint64_t)cin, (cuuint64_t)(padded_view ? wp : in.w),
                               (cuuint64_t)(tall ? (long long)in.n * hp : (padded_view ? hp : in.h)), (cuuint64_t)(tall ? 1 : in.n), 2};
  const cuuint64_t astr[4] = {(cuuint64_t)in.c_total * 2, (cuuint64_t)wp * in.c_total * 2, (cuuint64_t)hp * wp * in.c_total * 2,
                              (cuuint64_t)in.plane_stride * 2};
  const cuuint32_t abox[5] = {(cuuint32_t)kc, kTmaTW, (cuuint32_t)(kTmaTH + taps - 1), 1, 2};
  CUtensorMap ma, mb;
  if (int rc = encode(&ma, abase, 5, adims, astr, abox, kc * 2, "A")) return rc;
  // B: 3-D map (k, cout, plane); the lo plane must follow the hi plane in memory (one [2][...] allocation)
  const long long wdiff = reinterpret_cast<const char*>(p->w_lo) - reinterpret_cast<const char*>(p->w_hi);
  if (wdiff <= 0 || (wdiff & 15)) { set_error("conv3x3_tma: w_lo must follow w_hi (16-byte aligned distance)"); return -1; }
  const cuuint64_t bdims[3] = {(cuuint64_t)taps * taps * cin, (cuuint64_t)cout, 2};
  const cuuint64_t bstr[2] = {(cuuint64_t)taps * taps * cin * 2, (cuuint64_t)wdiff};
  const cuuint32_t bbox[3] = {(cuuint32_t)kc, (cuuint32_t)bn, 2};
  if (int rc = encode(&mb, const_cast<void*>(p->w_hi), 3, bdims, bstr, bbox, kc * 2, "B")) return rc;

  cudaStream_t st = as_stream(stream);
  if (workspace) {
    const long long items = (long long)g.n_img * g.tiles_x * g.tiles_y * g.n_blocks;
    const TmaSplit sp = tma_split_plan(items, g.chunks * taps, kc, bn, taps, p->sm_budget);
    if (sp.ksplit >= 2) {
      const long long slice = (long long)in.n * ho * wo * cout;
      if (workspace_bytes < sp.ksplit * slice * 4) {
        set_error("conv3x3_tma: workspace too small (%lld < %lld)", workspace_bytes, sp.ksplit * slice * 4); return -2;
      }
      // cluster form (CLSLAM_TMA_SPLITK_CLUSTER flag): at most 8 slices (portable cluster size), one CTA per work item
      int cks = 0;
      if (p->flags & CLSLAM_TMA_SPLITK_CLUSTER) {
        const int n_cs = g.chunks * taps;
        for (int c = 8; c >= 2; c >>= 1) {
          const int per = (n_cs + c - 1) / c;
          if (c <= sp.ksplit && items * c <= (p->sm_budget > 0 && p->sm_budget < tma_num_sms() ? p->sm_budget : tma_num_sms()) &&
              (long long)(c - 1) * per < n_cs) { cks = c; break; }
        }
      }
      if (cks > 0) {
        TmaConvArgs k = g;                    // full epilogue arguments: the cluster finishes the tile itself
        k.cluster_ks = cks; k.ksplit = cks; k.cs_per = (g.chunks * taps + cks - 1) / cks; k.div_ks = make_fastdiv(cks);
        k.part = nullptr; k.part_stride = 0; k.counters = nullptr;
        int rc;
        if (taps == 2 && tall) rc = launch_tma_conv<64, 128, 2, true, true>(ma, mb, k, st);
        else if (taps == 2) rc = launch_tma_conv<64, 128, 2, true>(ma, mb, k, st);
        else if (tall) rc = launch_tma_conv<64, 128, 3, true, true>(ma, mb, k, st);
        else if (bn == 128) rc = launch_tma_conv<64, 128, 3, true>(ma, mb, k, st);
        else rc = launch_tma_conv<64, 64, 3, true>(ma, mb, k, st);
        return rc;          // (a dgrad caller then runs the plain fold on p->y: `folded` stays false)
      }
      // fused reduction: the CTA holding the last slice of a tile reduces and runs the epilogue (needs the caller's counters;
      // the dgrad fold keeps its own reduction, and more tiles than counters fall back to the epilogue kernel)
      const bool fused = p->splitk_counters != nullptr && fold == nullptr && items <= CLSLAM_TMA_SPLITK_COUNTERS;
      TmaConvArgs k = g;
      if (!fused) {                         // the main kernel only stores raw partial sums
        k.scale = nullptr; k.shift = nullptr; k.res = PlaneView{}; k.gate = nullptr; k.y = nullptr; k.out = PlaneView{};
        k.act = CLSLAM_ACT_NONE;
      }
      k.counters = fused ? p->splitk_counters : nullptr;
      k.ksplit = sp.ksplit; k.cs_per = sp.cs_per; k.div_ks = make_fastdiv(sp.ksplit);
      k.part = reinterpret_cast<float*>(workspace); k.part_stride = slice;
      int rc;
      if (taps == 2 && tall) rc = launch_tma_conv<64, 128, 2, true, true>(ma, mb, k, st);
      else if (taps == 2) rc = launch_tma_conv<64, 128, 2, true>(ma, mb, k, st);
      else if (tall) rc = launch_tma_conv<64, 128, 3, true, true>(ma, mb, k, st);
      else if (bn == 128) rc = launch_tma_conv<64, 128, 3, true>(ma, mb, k, st);
      else rc = launch_tma_conv<64, 64, 3, true>(ma, mb, k, st);
      if (rc) return rc;
      if (fused) return 0;
      if (fold) {        // dgrad of a reflection-padded layer: the fold adds the slices itself
        if (folded) *folded = true;
        return launch_fold(k.part, sp.ksplit, slice, in.n, in.h, in.w, cout, *fold, st);
      }
      TmaSplitkEpiArgs e{};
      e.n_img = in.n; e.ho = ho; e.wo = wo; e.cout = cout; e.ksplit = sp.ksplit;
      e.part_stride = slice; e.part = k.part;
      e.act = g.act; e.scale = g.scale; e.shift = g.shift; e.res = g.res; e.gate = g.gate; e.gate_act = g.gate_act;
      e.y = g.y; e.out = g.out;
      const long long threads = slice / 8;
      int blocks = (int)((threads + 255) / 256);
      if (blocks > 148 * 8) blocks = 148 * 8;
      if (blocks < 1) blocks = 1;
      launch_k(tma_splitk_epilogue_kernel, dim3(blocks), dim3(256), 0, st, e);
      return finish_launch("conv3x3_tma (split-K epilogue)");
    }
  }
  if (taps == 4) {
    if (kc == 16 && bn == 64) return launch_tma_conv<16, 64, 4>(ma, mb, g, st);
    if (kc == 32 && bn == 64) return launch_tma_conv<32, 64, 4>(ma, mb, g, st);
    set_error("conv3x3_tma: the 4x4 form is built for 16 / 32 input channels and 64 output channels");
    return -1;
  }
  if (taps == 1) {        // EXPERIMENTAL (not yet run on hardware): 1x1 convolution, e.g. the stride-2 down-sampling branch of a
                          // BasicBlock read from the (p, q) = (0, 0) channel block of the space-to-depth planes
    if (kc == 64 && bn == 128) return launch_tma_conv<64, 128, 1>(ma, mb, g, st);
    set_error("conv3x3_tma: the 1x1 form is built for input channels %% 64 == 0 and output channels %% 128 == 0");
    return -1;
  }
  if (taps == 2) {
    if (kc == 64 && bn == 128 && tall) return launch_tma_conv<64, 128, 2, false, true>(ma, mb, g, st);
    if (kc == 64 && bn == 128) return launch_tma_conv<64, 128, 2>(ma, mb, g, st);
    set_error("conv3x3_tma: the 2x2 form is built for input channels %% 64 == 0 and output channels %% 128 == 0");
    return -1;
  }
  if (tall) return launch_tma_conv<64, 128, 3, false, true>(ma, mb, g, st);
#define CLSLAM_TMA_CASE(KC_, BN_) if (kc == KC_ && bn == BN_) return launch_tma_conv<KC_, BN_, 3>(ma, mb, g, st);
  CLSLAM_TMA_CASE(64, 128) CLSLAM_TMA_CASE(64, 64) CLSLAM_TMA_CASE(64, 32) CLSLAM_TMA_CASE(64, 16)
  CLSLAM_TMA_CASE(32, 128) CLSLAM_TMA_CASE(32, 64) CLSLAM_TMA_CASE(32, 32) CLSLAM_TMA_CASE(32, 16)
  CLSLAM_TMA_CASE(16, 64) CLSLAM_TMA_CASE(16, 32) CLSLAM_TMA_CASE(16, 16)
#undef CLSLAM_TMA_CASE
  set_error("conv3x3_tma: no kernel for channel chunk %d / N block %d", kc, bn);
  return -1;
}

static int launch_fold(const float* dxp, int ksplit, long long part_stride, int n, int h, int w, int c, const FoldArgs& f, cudaStream_t st) {
  const int up = f.up;
  if ((up != 1 && up != 2) || c % 4 || h % up || w % up || h < 3 || w < 3) { set_error("dgrad_fold: bad dims"); return -1; }
  if (f.disp_dz && up != 1) { set_error("dgrad_fold: dispconv term needs up == 1"); return -1; }
  PlaneView po{};
  if (f.dz_prev_planes) {
    const clslam_planes* q = f.dz_prev_planes;
    if (!q->base || q->n != n || q->h != h / up || q->w != w / up || q->c != c || q->c_total % 8 || q->c_off % 4) {
      set_error("dgrad_fold: planes mismatch"); return -1;
    }
    po = make_plane_view(q, 1, 0);
  }
  if (!f.dz_prev && !po.base) { set_error("dgrad_fold: no output"); return -1; }
  const size_t total = (size_t)n * (h / up) * (w / up) * (c / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  if (ksplit > 1) {
    launch_k(dgrad_fold_splitk_kernel, dim3(blocks), dim3(256), 0, st, dxp, ksplit, part_stride, n, h, w, c, up, f.y_prev, f.act_prev,
             f.disp_dz, f.disp_w, f.dz_prev, po);
    return finish_launch("dgrad_fold (split-K)");
  }
  launch_k(dgrad_fold_kernel, dim3(blocks), dim3(256), 0, st, dxp, n, h, w, c, up, f.y_prev, f.act_prev, f.disp_dz, f.disp_w, f.dz_prev, po);
  return finish_launch("dgrad_fold");
}

extern "C" {

int clslam_stage_images_s2d(const float* img_a, const float* img_b, const clslam_planes* out, int image_offset, int n, int h, int w,
                            void* stream) {
  if (int rc = check_planes(out, "stage_images_s2d")) return rc;
  const int cs = img_b ? 32 : 16;
  if (h % 2 || w % 2 || out->h != h / 2 || out->w != w / 2 || out->c != cs || image_offset < 0 || image_offset + n > out->n) {
    set_error("stage_images_s2d: output planes must be (>= %d, %d, %d, %d)", image_offset + n, h / 2, w / 2, cs); return -1;
  }
  const PlaneView v = make_plane_view(out, 1, 0);
  const size_t total = (size_t)n * (h / 2) * (w / 2) * 4;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(stage_s2d_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), img_a, img_b, v, n, h, w, image_offset);
  return finish_launch("stage_images_s2d");
}

int clslam_planes_s2d_from_f32(const float* x, const clslam_planes* out, int n, int h, int w, int c, void* stream) {
  if (int rc = check_planes(out, "planes_s2d_from_f32")) return rc;
  if (h % 2 || w % 2 || c % 4 || out->n != n || out->h != h / 2 || out->w != w / 2 || out->c != 4 * c) {
    set_error("planes_s2d_from_f32: output planes must be (%d, %d, %d, %d)", n, h / 2, w / 2, 4 * c); return -1;
  }
  const PlaneView v = make_plane_view(out, 1, 0);
  const size_t total = (size_t)n * h * w * (c / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(s2d_planes_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), x, v, n, h, w, c);
  return finish_launch("planes_s2d_from_f32");
}

int clslam_wt_flip(const void* items_dev, int n_items, int max_elems, const void* w_hi, long long w_plane_stride, void* wt_hi,
                   long long wt_plane_stride, void* stream) {
  if (n_items < 1 || !items_dev || !w_hi || !wt_hi) { set_error("wt_flip: bad arguments"); return -1; }
  int bx = (max_elems + 255) / 256;
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n_items);
  launch_k(wt_flip_kernel, dim3(grid), dim3(256), 0, as_stream(stream), reinterpret_cast<const WtFlipItem*>(items_dev),
                                                      reinterpret_cast<const uint16_t*>(w_hi), w_plane_stride,
                                                      reinterpret_cast<uint16_t*>(wt_hi), wt_plane_stride);
  return finish_launch("wt_flip");
}

int clslam_dgrad_fold(const float* dxp, int n, int h, int w, int c, int up, const float* y_prev, int act_prev,
                      const float* disp_dz, const float* disp_w, float* dz_prev, const clslam_planes* dz_prev_planes,
                      void* stream) {
  const FoldArgs f{up, y_prev, act_prev, disp_dz, disp_w, dz_prev, dz_prev_planes};
  return launch_fold(dxp, 1, 0, n, h, w, c, f, as_stream(stream));
}

int clslam_dgrad3x3_tma_ws(const clslam_conv_tma* p, int up, const float* y_prev, int act_prev, const float* disp_dz,
                           const float* disp_w, float* dz_prev, const clslam_planes* dz_prev_planes, void* workspace,
                           long long workspace_bytes, void* stream) {
  if (!p || p->pad_mode != CLSLAM_PAD_FULL || !p->y || p->out || p->residual || p->gate || p->scale || p->shift ||
      p->act != CLSLAM_ACT_NONE) {
    set_error("dgrad3x3_tma: needs a plain PAD_FULL launch with the fp32 scratch output y"); return -1;
  }
  const FoldArgs f{up, y_prev, act_prev, disp_dz, disp_w, dz_prev, dz_prev_planes};
  bool folded = false;
  if (int rc = conv3x3_tma_impl(p, workspace, workspace_bytes, stream, &f, &folded)) return rc;
  if (folded) return 0;          // split-K: the fold read the partial sums directly
  return launch_fold(p->y, 1, 0, p->in.n, p->in.h, p->in.w, p->cout, f, as_stream(stream));
}

int clslam_debug_tma_flags(int flags) { g_tma_debug = flags; return 0; }

int clslam_debug_tma_stamps(unsigned long long* out32) {
#ifdef CLSLAM_TMA_DEBUG
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out32, g_tma_stamps, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : -1;
#else
  for (int i = 0; i < 32; ++i) out32[i] = 0;
  set_error("clslam_debug_tma_stamps: the library was built without -DCLSLAM_TMA_DEBUG");
  return -1;
#endif
}

int clslam_planes_from_f32(const float* x, const clslam_planes* out, int out_up, int out_halo, void* stream) {
  if (int rc = check_planes(out, "planes_from_f32")) return rc;
  if (out_up != 1 && out_up != 2) { set_error("planes_from_f32: up"); return -1; }
  const PlaneView v = make_plane_view(out, out_up, out_halo);
  const size_t total = (size_t)out->n * v.h_lo * v.w_lo * (out->c / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  launch_k(planes_from_f32_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), x, v, out->n, v.h_lo, v.w_lo, out->c);
  return finish_launch("planes_from_f32");
}

}  // extern "C"
