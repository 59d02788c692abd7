// tcgen05 forward kernel specialised for 3x3 / stride-1 / pad-1 convolutions (zero or reflection
// padding, plain or nearest-up + concat input) with Cin a multiple of 64 and weights pre-split into
// bf16 hi / lo planes ([Cout][9*Cin], K-contiguous).
//
// Compared with the generic gather kernel (gemm_tc.cuh):
//   * every input element is read from global memory, split into bf16 hi+lo and written to shared
//     memory ONCE per CTA and 64-channel chunk (a (8+2) x (16+2) pixel patch); the nine tap operands of
//     the chunk are assembled by 16-byte shared->shared copies into the swizzled UMMA tile;
//   * the weight tiles are cp.async (LDGSTS) copies of the pre-split planes straight into their
//     swizzled slots, issued two k blocks ahead (3 B stages) so global latency never sits on the
//     critical path;
//   * warp 8 only issues tcgen05.mma / tcgen05.commit; warps 0-7 produce and run the epilogue.
//
// M tile = 8 x 16 output pixels of one image (TMEM lane l <-> pixel (l / 16, l % 16)).
#pragma once
#include "conv_ops.cuh"
#include "gemm_tc.cuh"

namespace clslam {

constexpr int kP3TH = 16, kP3TW = 8;                       // output tile (8 wide: see the A descriptor below)
constexpr int kP3PH = kP3TH + 2, kP3PW = kP3TW + 2;        // input patch
constexpr int kP3Pix = kP3PH * kP3PW;                      // 180
constexpr int kP3PlaneChunk = kP3Pix * 16;                 // bytes of one 8-channel chunk of one plane: [pixel][16 B]
constexpr int kP3PatchBytes = 8 * kP3PlaneChunk;           // one bf16 plane (hi or lo), 64 channels: [chunk8][pixel][16 B]
constexpr int kP3PatchVec = (kP3Pix * 16 + kTcThreads - 1) / kTcThreads;   // float4 gathers per thread and chunk (12)
constexpr int kP3Patches = 2, kP3BStages = 3;

template <int BN>
struct P3Smem {
  static constexpr int kBBytes = BN * 128;                 // one plane of a B tile (K-major, SWIZZLE_128B)
  static constexpr int kBStage = 2 * kBBytes;
  static constexpr int kPatch = 2 * kP3PatchBytes;         // hi + lo
  static constexpr int kParamBytes = 2 * BN * 4;
  static constexpr int kTotal = kP3BStages * kBStage + kP3Patches * kPatch + kParamBytes + 1024 + 128;
};

struct P3Geom {
  int tiles_x, tiles_y;      // M tiles per image
  int n_chunks;              // ci / 64
  int chunks_per_split;
};

// one float4 (4 channels) of the conv input at padded patch position; zero outside / for overhanging tiles
__device__ __forceinline__ float4 p3_gather(const ConvIn& c, int n, int ih, int iw, int ch) {
  if (c.pad_mode == CLSLAM_PAD_REFLECT) {
    if (ih > c.hi || iw > c.wi) return f4zero();           // beyond the reflected border: only feeds discarded outputs
    ih = reflect_idx(ih, c.hi);
    iw = reflect_idx(iw, c.wi);
  } else if (ih < 0 || ih >= c.hi || iw < 0 || iw >= c.wi) {
    return f4zero();
  }
  if (c.in_mode == CLSLAM_IN_UPCAT) {
    if (ch < c.c1) {
      const int sh = (2 * c.h1 == c.hi) ? (ih >> 1) : (int)(((long long)ih * c.h1) / c.hi);
      const int sw = (2 * c.w1 == c.wi) ? (iw >> 1) : (int)(((long long)iw * c.w1) / c.wi);
      return ld4(c.x + (((size_t)n * c.h1 + sh) * c.w1 + sw) * c.c1 + ch);
    }
    return ld4(c.x2 + (((size_t)n * c.hi + ih) * c.wi + iw) * (c.ci - c.c1) + (ch - c.c1));
  }
  return ld4(c.x + (((size_t)n * c.hi + ih) * c.wi + iw) * c.ci + ch);
}

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, bool valid) {
  const int bytes = valid ? 16 : 0;                        // src-size 0 -> the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(bytes) : "memory");
}

// A operand straight out of the patch: K-major, NO swizzle ("interleaved" canonical layout,
// cute::UMMA::LayoutType::SWIZZLE_NONE = ((8,n),2):((1,SBO),LBO) in 16-byte units):
//   8 rows (= 8 consecutive pixels of one output row) x 16 B are contiguous, the next 8-row group
//   (next output row) is SBO = patch_width * 16 B away, the next 8 channels LBO = one plane chunk away.
__device__ __forceinline__ uint64_t umma_desc_k_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

template <int BN, class Epi>
__global__ void __launch_bounds__(kTcBlock)
conv3x3_tc_kernel(ConvIn c, const uint16_t* __restrict__ w_hi, const uint16_t* __restrict__ w_lo, Epi epi, int co, P3Geom g,
                  float* __restrict__ ws) {
  pdl_wait();
  pdl_trigger();
  using SM = P3Smem<BN>;
  constexpr int kTmemCols = BN < 32 ? 32 : BN;
  constexpr int kBVec = (BN * 8 * 2 + kTcThreads - 1) / kTcThreads;      // 16-byte copies per thread and k block (hi + lo)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* b_tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* patches = b_tiles + kP3BStages * SM::kBStage;
  float* eparams = reinterpret_cast<float*>(patches + kP3Patches * SM::kPatch);
  uint64_t* b_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(eparams) + SM::kParamBytes);
  uint64_t* b_empty = b_full + kP3BStages;
  uint64_t* patch_empty = b_empty + kP3BStages;
  uint64_t* done = patch_empty + kP3Patches;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool producer = warp < kTcThreads / 32;
  const int tiles_per_img = g.tiles_x * g.tiles_y;
  const int img = blockIdx.x / tiles_per_img;
  const int trem = blockIdx.x - img * tiles_per_img;
  const int oh0 = (trem / g.tiles_x) * kP3TH, ow0 = (trem % g.tiles_x) * kP3TW;
  const int n0 = blockIdx.y * BN;
  const int chunk_begin = blockIdx.z * g.chunks_per_split;
  const int chunk_end = min(g.n_chunks, chunk_begin + g.chunks_per_split);
  const int n_kb = (chunk_end - chunk_begin) * 9;
  const int K = 9 * c.ci;
  const int M = c.n * c.ho * c.wo, N = co;

  TC_STAMP(0);
  if (tid == 0) {
    for (int s = 0; s < kP3BStages; ++s) { mbar_init(&b_full[s], kTcThreads / 32); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < kP3Patches; ++s) mbar_init(&patch_empty[s], 1);
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!producer) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    epi.tc_prologue(eparams, n0, BN, N, tid);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  constexpr uint32_t idesc = umma_idesc_bf16(BN);
  TC_STAMP(1);

  if (producer) {
    // ---- patch staging: gather -> registers (fetch) -> split -> shared memory (commit) -------------
    float4 pv[kP3PatchVec];
    auto patch_fetch = [&](int chunk) {
#pragma unroll
      for (int i = 0; i < kP3PatchVec; ++i) {
        const int f = tid + i * kTcThreads;
        const int pix = f >> 4;
        if (pix < kP3Pix) {
          const int py = pix / kP3PW, px = pix - py * kP3PW;
          pv[i] = p3_gather(c, img, oh0 - 1 + py, ow0 - 1 + px, chunk * 64 + (f & 15) * 4);
        }
      }
    };
    auto patch_commit = [&](uint8_t* p_hi) {
#pragma unroll
      for (int i = 0; i < kP3PatchVec; ++i) {
        const int f = tid + i * kTcThreads;
        const int pix = f >> 4, c4 = f & 15;
        if (pix < kP3Pix)       // [chunk8 = c4 / 2][pixel][16 B], this float4 fills half (8 B) of a 16-byte slot
          store_split4(p_hi, p_hi + kP3PatchBytes, (uint32_t)((c4 >> 1) * kP3PlaneChunk + pix * 16 + (c4 & 1) * 8), pv[i].x,
                       pv[i].y, pv[i].z, pv[i].w);
      }
    };
    // ---- weight tile of k block kb: cp.async copies of the pre-split planes into their swizzled slots
    auto b_issue = [&](int kb) {
      const int chunk = chunk_begin + kb / 9, tap = kb % 9;
      const size_t kofs = (size_t)tap * c.ci + (size_t)chunk * 64;
      uint8_t* stage = b_tiles + (kb % kP3BStages) * SM::kBStage;
#pragma unroll
      for (int i = 0; i < kBVec; ++i) {
        const int f = tid + i * kTcThreads;              // [plane][row][chunk16]
        if (f >= BN * 16) continue;
        const int plane = f / (BN * 8), r = (f / 8) % BN, c16 = f & 7;
        const bool ok = n0 + r < N;
        const uint16_t* src = (plane ? w_lo : w_hi) + (size_t)(ok ? n0 + r : 0) * K + kofs + c16 * 8;
        cp_async16(smem_u32(stage + plane * SM::kBBytes + sw128_offset(r, c16 * 8)), src, ok);
      }
    };

    if (n_kb > 0) {
      patch_fetch(chunk_begin);
      b_issue(0);
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (n_kb > 1) b_issue(1);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kb = 0; kb < n_kb; ++kb) {
      const int tap = kb % 9, cj = kb / 9;
      if (tap == 0) {                 // new 64-channel chunk -> patch buffer cj % 2 (free once chunk cj-2's MMAs are done)
        if (cj >= kP3Patches) mbar_wait(&patch_empty[cj % kP3Patches], (uint32_t)((cj / kP3Patches - 1) & 1));
        patch_commit(patches + (cj % kP3Patches) * SM::kPatch);
        if (kb == 0) TC_STAMP(2);
      }
      if (kb + 2 < n_kb) {            // weights two k blocks ahead
        if (kb + 2 >= kP3BStages) mbar_wait(&b_empty[(kb + 2) % kP3BStages], (uint32_t)(((kb + 2) / kP3BStages - 1) & 1));
        b_issue(kb + 2);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 2;" ::: "memory");      // this thread's copies of B(kb) have landed
      fence_proxy_async_smem();                                   // ... and (tap 0) its part of the patch is visible
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[kb % kP3BStages]);
      if (kb == 0) TC_STAMP(3);
      if (tap == 0 && kb + 9 < n_kb) patch_fetch(chunk_begin + cj + 1);     // next chunk's gathers fly during this chunk's taps
    }
    TC_STAMP(4);
    if (n_kb > 0) mbar_wait(done, 0);
    TC_STAMP(5);
  } else {
    // ===== warp 8: MMA issuer =====
    for (int kb = 0; kb < n_kb; ++kb) {
      const int sb = kb % kP3BStages, tap = kb % 9, cj = kb / 9;
      mbar_wait(&b_full[sb], (uint32_t)((kb / kP3BStages) & 1));
      tc_fence_after();
      if (lane == 0) {
        const int r = tap / 3, s = tap - r * 3;
        uint8_t* p_hi = patches + (cj % kP3Patches) * SM::kPatch + (r * kP3PW + s) * 16;
        uint8_t* b_hi = b_tiles + sb * SM::kBStage;
        const uint64_t dah = umma_desc_k_nosw(smem_u32(p_hi), kP3PlaneChunk, kP3PW * 16);
        const uint64_t dal = umma_desc_k_nosw(smem_u32(p_hi + kP3PatchBytes), kP3PlaneChunk, kP3PW * 16);
        const uint64_t dbh = umma_desc_k_sw128(smem_u32(b_hi)), dbl = umma_desc_k_sw128(smem_u32(b_hi + SM::kBBytes));
#pragma unroll
        for (int j = 0; j < kTcBK / 16; ++j) {
          const uint64_t adv_a = (uint64_t)((j * 2 * kP3PlaneChunk) >> 4);     // 16 channels = two 8-channel chunks
          const uint64_t adv_b = (uint64_t)((j * 32) >> 4);
          umma_bf16(tmem_d, dal + adv_a, dbh + adv_b, idesc, (kb > 0 || j > 0) ? 1u : 0u);
          umma_bf16(tmem_d, dah + adv_a, dbl + adv_b, idesc, 1u);
          umma_bf16(tmem_d, dah + adv_a, dbh + adv_b, idesc, 1u);
        }
        umma_commit(&b_empty[sb]);
        if (tap == 8) umma_commit(&patch_empty[cj % kP3Patches]);
        if (kb == n_kb - 1) umma_commit(done);
      }
      __syncwarp();
    }
  }
  tc_fence_after();
  constexpr int kColsPerWarp = (BN >= 64) ? BN / 2 : BN;
  const int quarter = warp & 3, half = (warp >> 2) & 1;
  const bool active = producer && ((BN >= 64) || half == 0);
  if (active) {
    const int l = quarter * 32 + lane;
    const int oh = oh0 + (l >> 3), ow = ow0 + (l & 7);
    const bool row_ok = oh < c.ho && ow < c.wo;
    const int m = (img * c.ho + oh) * c.wo + ow;
    const int col0 = (BN >= 64) ? half * kColsPerWarp : 0;
    float* part = (gridDim.z > 1) ? ws + (size_t)blockIdx.z * M * N : nullptr;
#pragma unroll 1
    for (int cc = 0; cc < kColsPerWarp; cc += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(col0 + cc);
      if (n_kb > 0) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = 0u;
      }
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int n = n0 + col0 + cc + j;
          if (n < N) {
            const float v[4] = {__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])};
            if (part) *reinterpret_cast<float4*>(&part[(size_t)m * N + n]) = make_float4(v[0], v[1], v[2], v[3]);
            else epi.tc_store(m, n, v, eparams, col0 + cc + j, BN);
          }
        }
      }
    }
  }
  TC_STAMP(6);
  tc_fence_before();
  __syncthreads();
  if (!producer) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
  }
}

// fp32 -> bf16 hi / lo planes (weights are split once per update instead of once per tile)
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, long long n4) {
  pdl_wait();
  pdl_trigger();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ld4(src + i * 4);
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
    uint2 hv, lv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h0); hv.y = *reinterpret_cast<const uint32_t*>(&h1);
    lv.x = *reinterpret_cast<const uint32_t*>(&l0); lv.y = *reinterpret_cast<const uint32_t*>(&l1);
    *reinterpret_cast<uint2*>(hi + i * 4) = hv;
    *reinterpret_cast<uint2*>(lo + i * 4) = lv;
  }
}

}  // namespace clslam
