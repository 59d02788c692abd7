// tcgen05 / TMEM implicit-GEMM tile kernel (sm_100a) with 3xBF16 split operands.
//
// C[M][N] (fp32, TMEM) = sum_k A[m][k] * B[k][n], A and B are fp32 in global memory and are
// gathered by the operand policies (reflect padding, nearest-up + concat, dgrad tap sums, ...),
// split on the fly into bf16 hi + lo parts and stored into shared memory in the canonical UMMA
// K-major SWIZZLE_128B layout.  One elected thread issues, per 16-wide k step, the three products
//     D += A_lo.B_hi + A_hi.B_lo + A_hi.B_hi            (tcgen05.mma.cta_group::1.kind::f16)
// with fp32 accumulation in tensor memory, which keeps ~16 mantissa bits per operand (single-pass
// BF16 / TF32 miss the 1e-4 pose tolerance).  kTcStages shared-memory stages:
// while the tensor core works on stage s, all 256 threads gather and convert the next stage;
// tcgen05.commit on an mbarrier releases a stage.  Epilogue: tcgen05.ld (32 lanes x 16 columns per
// warp and instruction) -> registers -> the epilogue functors (or raw split-K partials).
//
// The gather loops are deliberately NOT fully unrolled and the per-row decode lives in a small
// shared-memory table: the kernels stay a few KB of SASS, which matters because every launch of
// the adaptation step starts with a cold instruction cache (L2 is flushed between steps).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_prims.cuh"

namespace clslam {

constexpr int kTcBK = 64;       // k per stage = one 128-byte swizzle row of bf16
constexpr int kTcThreads = 256;            // producer (staging / epilogue) threads: warps 0..7
constexpr int kTcBlock = kTcThreads + 32;  // + warp 8: the MMA issuer
constexpr int kTcStages = 2;

// byte offset of element (row, k) (k multiple of 4) inside a [rows][64] bf16 K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + ((k & 7) << 1));
}

// split 4 fp32 into bf16 hi / lo and store both 8-byte groups
__device__ __forceinline__ void store_split4(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, float a, float b, float c, float d) {
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(a, b), h1 = __floats2bfloat162_rn(c, d);
  const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
  const __nv_bfloat162 l0 = __floats2bfloat162_rn(a - f0.x, b - f0.y), l1 = __floats2bfloat162_rn(c - f1.x, d - f1.y);
  uint2 hv, lv;
  hv.x = *reinterpret_cast<const uint32_t*>(&h0); hv.y = *reinterpret_cast<const uint32_t*>(&h1);
  lv.x = *reinterpret_cast<const uint32_t*>(&l0); lv.y = *reinterpret_cast<const uint32_t*>(&l1);
  *reinterpret_cast<uint2*>(hi_tile + off) = hv;
  *reinterpret_cast<uint2*>(lo_tile + off) = lv;
}

// ---- operand staging ----------------------------------------------------------------------
// Operand policy interface (see conv.cu):
//   KContig (memory is contiguous along k for a fixed row):
//     Row    decode(int row)          once per row and CTA (kept in shared memory; out of range -> invalid)
//     KState kdecode(int k)           once per thread and k block (all rows of a thread share k)
//     float4 at(Row, KState)          elements (row, k..k+3)
//   RContig (memory is contiguous along the rows for a fixed k):
//     RowState rdecode(int row)       once per 4-row group and CTA (shared memory)
//     KState   kdecode(int k)         4 per thread and k block
//     float4   at4(KState, RowState)  elements (row..row+3, k)
template <class Op, bool KContig = Op::kKContig>
struct TcRowType { using type = typename Op::Row; };
template <class Op>
struct TcRowType<Op, false> { using type = typename Op::RowState; };

template <int R, class Op, bool KContig = Op::kKContig>
struct TcStager;

// Register-prefetch stagers: fetch() issues every global load of the NEXT k block (they stay in
// flight while the tensor core works on the current one), commit() converts and stores them.
template <int R, class Op>
struct TcStager<R, Op, true> {
  using RowT = typename Op::Row;
  static constexpr int kTableEntries = R;
  static constexpr int kVec = (R * 16 + kTcThreads - 1) / kTcThreads;      // float4 per thread and stage
  RowT rows[kVec];
  float4 v[kVec];
  __device__ static void build_table(const Op& op, RowT* tab, int origin, int tid) {
    for (int r = tid; r < R; r += kTcThreads) tab[r] = op.decode(origin + r);
  }
  __device__ void init(const RowT* tab, int tid) {
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int r = (tid >> 4) + i * (kTcThreads / 16);
      rows[i] = tab[r < R ? r : 0];
    }
  }
  __device__ void fetch(const Op& op, int k0, int k_end, int tid) {
    const int k = k0 + (tid & 15) * 4;          // identical for all rows of this thread
    const bool kv = k < k_end;
    const typename Op::KState ks = op.kdecode(kv ? k : k0);
#pragma unroll
    for (int i = 0; i < kVec; ++i) v[i] = kv ? op.at(rows[i], ks) : f4zero();
  }
  __device__ void commit(uint8_t* hi, uint8_t* lo, int tid) const {
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int r = (tid >> 4) + i * (kTcThreads / 16);
      if (r < R) store_split4(hi, lo, sw128_offset(r, (tid & 15) * 4), v[i].x, v[i].y, v[i].z, v[i].w);
    }
  }
};

template <int R, class Op>
struct TcStager<R, Op, false> {
  using RowT = typename Op::RowState;
  static constexpr int kTableEntries = R / 4;
  static constexpr int kVec = (R / 4 * 16 + kTcThreads - 1) / kTcThreads;  // 4x4 blocks per thread and stage
  RowT rs[kVec];
  float4 v[kVec][4];
  __device__ static void build_table(const Op& op, RowT* tab, int origin, int tid) {
    for (int r = tid; r < R / 4; r += kTcThreads) tab[r] = op.rdecode(origin + r * 4);
  }
  __device__ void init(const RowT* tab, int tid) {
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int rq = (tid >> 4) + i * (kTcThreads / 16);
      rs[i] = tab[rq < R / 4 ? rq : 0];
    }
  }
  __device__ void fetch(const Op& op, int k0, int k_end, int tid) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + (tid & 15) * 4 + j;
      const bool kv = k < k_end;
      const typename Op::KState ks = op.kdecode(kv ? k : k0);
#pragma unroll
      for (int i = 0; i < kVec; ++i) v[i][j] = kv ? op.at4(ks, rs[i]) : f4zero();
    }
  }
  __device__ void commit(uint8_t* hi, uint8_t* lo, int tid) const {
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int rq = (tid >> 4) + i * (kTcThreads / 16);
      if (rq >= R / 4) continue;
      const int kq = tid & 15;
      store_split4(hi, lo, sw128_offset(rq * 4 + 0, kq * 4), v[i][0].x, v[i][1].x, v[i][2].x, v[i][3].x);
      store_split4(hi, lo, sw128_offset(rq * 4 + 1, kq * 4), v[i][0].y, v[i][1].y, v[i][2].y, v[i][3].y);
      store_split4(hi, lo, sw128_offset(rq * 4 + 2, kq * 4), v[i][0].z, v[i][1].z, v[i][2].z, v[i][3].z);
      store_split4(hi, lo, sw128_offset(rq * 4 + 3, kq * 4), v[i][0].w, v[i][1].w, v[i][2].w, v[i][3].w);
    }
  }
};

template <int BN, class AOp, class BOp>
struct TcSmem {
  using ARow = typename TcRowType<AOp>::type;
  using BRow = typename TcRowType<BOp>::type;
  static constexpr int kABytes = kTcBM * kTcBK * 2;        // one bf16 A tile (hi or lo)
  static constexpr int kBBytes = BN * kTcBK * 2;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kTabABytes = (int)((sizeof(ARow) * TcStager<kTcBM, AOp>::kTableEntries + 15) / 16 * 16);
  static constexpr int kTabBBytes = (int)((sizeof(BRow) * TcStager<BN, BOp>::kTableEntries + 15) / 16 * 16);
  static constexpr int kParamBytes = 2 * BN * 4;          // epilogue per-column parameters (scale, shift)
  static constexpr int kTotal = kTcStages * kStageBytes + kTabABytes + kTabBBytes + kParamBytes + 1024 /*alignment slack*/ +
                                128 /*barriers + tmem slot*/;
  static_assert((2 * kTcStages + 1) * 8 + 4 <= 128, "barrier area");
};

// cycle stamps of CTA (0,0,0), thread 0 (debug aid, read back with clslam_debug_tc_stamps)
__device__ unsigned long long g_tc_stamps[16];
#define TC_STAMP(i) do { if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_tc_stamps[i] = clock64(); } while (0)

template <int BN, class AOp, class BOp, class Epi>
__global__ void __launch_bounds__(kTcBlock)
gemm_tc_kernel(AOp aop, BOp bop, Epi epi, int M, int N, int K, int k_chunk, float* __restrict__ ws) {
  pdl_wait();
  pdl_trigger();
  using SM = TcSmem<BN, AOp, BOp>;
  using StA = TcStager<kTcBM, AOp>;
  using StB = TcStager<BN, BOp>;
  constexpr int kTmemCols = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* after = tiles + kTcStages * SM::kStageBytes;
  typename SM::ARow* tab_a = reinterpret_cast<typename SM::ARow*>(after);
  typename SM::BRow* tab_b = reinterpret_cast<typename SM::BRow*>(after + SM::kTabABytes);
  float* eparams = reinterpret_cast<float*>(after + SM::kTabABytes + SM::kTabBBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(after + SM::kTabABytes + SM::kTabBBytes + SM::kParamBytes);
  uint64_t* empty = full + kTcStages;
  uint64_t* done = empty + kTcStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool producer = warp < kTcThreads / 32;
  const int m0 = blockIdx.x * kTcBM, n0 = blockIdx.y * BN;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);
  const int n_kb = (k_end - k_begin + kTcBK - 1) / kTcBK;

  TC_STAMP(0);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&full[s], kTcThreads / 32); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!producer) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    StA::build_table(aop, tab_a, m0, tid);
    StB::build_table(bop, tab_b, n0, tid);
    epi.tc_prologue(eparams, n0, BN, N, tid);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  constexpr uint32_t idesc = umma_idesc_bf16(BN);
  TC_STAMP(1);

  if (producer) {
    // ===== producers: gather + split + swizzled store, one k block ahead of the tensor core =====
    StA sa;
    StB sb;
    sa.init(tab_a, tid);
    sb.init(tab_b, tid);
    if (n_kb > 0) { sa.fetch(aop, k_begin, k_end, tid); sb.fetch(bop, k_begin, k_end, tid); }
    for (int kb = 0; kb < n_kb; ++kb) {
      const int s = kb % kTcStages;
      uint8_t* a_hi = tiles + s * SM::kStageBytes;
      uint8_t* a_lo = a_hi + SM::kABytes;
      uint8_t* b_hi = a_lo + SM::kABytes;
      uint8_t* b_lo = b_hi + SM::kBBytes;
      if (kb >= kTcStages) mbar_wait(&empty[s], (uint32_t)((kb / kTcStages - 1) & 1));   // MMAs that read this stage are done
      sa.commit(a_hi, a_lo, tid);
      sb.commit(b_hi, b_lo, tid);
      fence_proxy_async_smem();             // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      if (kb == 0) TC_STAMP(2);
      if (kb + 1 < n_kb) {                  // loads of the next k block fly while the MMAs run
        sa.fetch(aop, k_begin + (kb + 1) * kTcBK, k_end, tid);
        sb.fetch(bop, k_begin + (kb + 1) * kTcBK, k_end, tid);
      }
    }
    TC_STAMP(4);
    if (n_kb > 0) mbar_wait(done, 0);
  } else {
    // ===== warp 8: MMA issuer =====
    for (int kb = 0; kb < n_kb; ++kb) {
      const int s = kb % kTcStages;
      mbar_wait(&full[s], (uint32_t)((kb / kTcStages) & 1));
      tc_fence_after();
      if (lane == 0) {
        uint8_t* a_hi = tiles + s * SM::kStageBytes;
        uint8_t* a_lo = a_hi + SM::kABytes;
        uint8_t* b_hi = a_lo + SM::kABytes;
        uint8_t* b_lo = b_hi + SM::kBBytes;
        const uint64_t dah = umma_desc_k_sw128(smem_u32(a_hi)), dal = umma_desc_k_sw128(smem_u32(a_lo));
        const uint64_t dbh = umma_desc_k_sw128(smem_u32(b_hi)), dbl = umma_desc_k_sw128(smem_u32(b_lo));
#pragma unroll
        for (int j = 0; j < kTcBK / 16; ++j) {
          const uint64_t adv = (uint64_t)((j * 32) >> 4);       // 16 bf16 = 32 bytes inside the swizzle row
          umma_bf16(tmem_d, dal + adv, dbh + adv, idesc, (kb > 0 || j > 0) ? 1u : 0u);
          umma_bf16(tmem_d, dah + adv, dbl + adv, idesc, 1u);
          umma_bf16(tmem_d, dah + adv, dbh + adv, idesc, 1u);
        }
        umma_commit(&empty[s]);
        if (kb == n_kb - 1) umma_commit(done);
      }
      __syncwarp();
    }
  }
  // ---- epilogue (producer warps) ----------------------------------------------------------------
  tc_fence_after();
  TC_STAMP(5);
  constexpr int kColsPerWarp = (BN >= 64) ? BN / 2 : BN;          // 8 warps: 4 lane quarters x 2 column halves
  const int quarter = warp & 3, half = (warp >> 2) & 1;
  const bool active = producer && ((BN >= 64) || half == 0);
  if (active) {
    const int m = m0 + quarter * 32 + lane;
    const int col0 = (BN >= 64) ? half * kColsPerWarp : 0;
    float* part = (gridDim.z > 1) ? ws + (size_t)blockIdx.z * M * N : nullptr;
#pragma unroll 1
    for (int c = 0; c < kColsPerWarp; c += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(col0 + c);
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
      if (m < M) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int n = n0 + col0 + c + j;
          if (n < N) {
            const float v[4] = {__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])};
            if (part) *reinterpret_cast<float4*>(&part[(size_t)m * N + n]) = make_float4(v[0], v[1], v[2], v[3]);
            else epi.tc_store(m, n, v, eparams, col0 + c + j, BN);
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

}  // namespace clslam
