// fp32 CUDA-core implicit-GEMM tile kernel shared by conv forward / dgrad / wgrad.
//
// C[M][N] = sum_k A[m][k] * B[k][n].  Operand policies gather A and B from global memory
// (float4 granularity) into registers ("fetch") and then into K-major shared tiles
// As[BK][BM], Bs[BK][BN] ("commit"); the next tile is fetched while the current one is
// multiplied.  blockIdx.z splits the reduction: partial tiles go to `ws` and a second kernel
// (splitk_reduce_kernel) sums them in a fixed order and applies the epilogue, so results are
// run-to-run deterministic (the reference's CPU path is).
//
// This is the exact-fp32 baseline path; the tensor-core (tcgen05) path plugs in per layer.
#pragma once
#include "common.cuh"

namespace clslam {

constexpr int kBK = 16;
constexpr int kGemmThreads = 256;

// ---- operand staging helpers --------------------------------------------------------------
// "K-contiguous" operand: element (row, k..k+3) is one float4 in memory; R rows x kBK k's.
template <int R>
struct KContigMap {
  static constexpr int kVec = (R * 4 + kGemmThreads - 1) / kGemmThreads;
  __device__ static bool valid(int tid, int v) { return tid + v * kGemmThreads < R * 4; }
  __device__ static int row(int tid, int v) { return (tid + v * kGemmThreads) >> 2; }
  __device__ static int k4(int tid, int v) { return (tid + v * kGemmThreads) & 3; }
  __device__ static void commit(float* S, int ld, int tid, const float4 (&r)[kVec]) {
#pragma unroll
    for (int v = 0; v < kVec; ++v) {
      if (!valid(tid, v)) continue;
      const int rw = row(tid, v), kk = k4(tid, v) * 4;
      S[(kk + 0) * ld + rw] = r[v].x;
      S[(kk + 1) * ld + rw] = r[v].y;
      S[(kk + 2) * ld + rw] = r[v].z;
      S[(kk + 3) * ld + rw] = r[v].w;
    }
  }
};
// "row-contiguous" operand: element (k, row..row+3) is one float4 in memory.
template <int R>
struct RContigMap {
  static constexpr int kPerK = R / 4;
  static constexpr int kVec = (kBK * kPerK + kGemmThreads - 1) / kGemmThreads;
  __device__ static bool valid(int tid, int v) { return tid + v * kGemmThreads < kBK * kPerK; }
  __device__ static int kk(int tid, int v) { return (tid + v * kGemmThreads) / kPerK; }
  __device__ static int r4(int tid, int v) { return ((tid + v * kGemmThreads) % kPerK) * 4; }
  __device__ static void commit(float* S, int ld, int tid, const float4 (&r)[kVec]) {
#pragma unroll
    for (int v = 0; v < kVec; ++v) {
      if (!valid(tid, v)) continue;
      *reinterpret_cast<float4*>(&S[kk(tid, v) * ld + r4(tid, v)]) = r[v];
    }
  }
};

// ---- the tile kernel ----------------------------------------------------------------------
// AOp / BOp interface (R = BM for A, BN for B):
//   template<int R> struct Frag;                       per-thread state incl. fetched registers
//   template<int R> __device__ void init(Frag<R>&, int origin, int tid) const;
//   template<int R> __device__ void fetch(Frag<R>&, int k0, int k_end, int tid) const;
//   template<int R> __device__ void commit(const Frag<R>&, float* S, int ld, int tid) const;
// Epi interface:
//   static constexpr bool kBias;  __device__ void store(int m, int n, const float (&v)[4]) const;
//   __device__ void store_bias(int m, float v) const;   (kBias only)
template <int BM, int BN, int TM, class AOp, class BOp, class Epi>
__global__ void __launch_bounds__(kGemmThreads)
gemm_tile_kernel(AOp aop, BOp bop, Epi epi, int M, int N, int K, int k_chunk, float* __restrict__ ws) {
  pdl_wait();
  pdl_trigger();
  constexpr int TN = 4;
  constexpr int TX = BN / TN;
  static_assert((BM / TM) * TX == kGemmThreads, "tile / thread mismatch");
  constexpr int LDA = BM + 4, LDB = BN + 4;
  __shared__ __align__(16) float As[kBK * LDA];
  __shared__ __align__(16) float Bs[kBK * LDB];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float bsum[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) bsum[i] = 0.f;

  typename AOp::template Frag<BM> fa;
  typename BOp::template Frag<BN> fb;
  aop.template init<BM>(fa, m0, tid);
  bop.template init<BN>(fb, n0, tid);
  if (k_begin < k_end) {
    aop.template fetch<BM>(fa, k_begin, k_end, tid);
    bop.template fetch<BN>(fb, k_begin, k_end, tid);
  }
  for (int k0 = k_begin; k0 < k_end; k0 += kBK) {
    aop.template commit<BM>(fa, As, LDA, tid);
    bop.template commit<BN>(fb, Bs, LDB, tid);
    __syncthreads();
    if (k0 + kBK < k_end) {
      aop.template fetch<BM>(fa, k0 + kBK, k_end, tid);
      bop.template fetch<BN>(fb, k0 + kBK, k_end, tid);
    }
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[TM], b[TN];
      if constexpr (TM % 4 == 0) {
#pragma unroll
        for (int i = 0; i < TM; i += 4) {
          const float4 t = *reinterpret_cast<const float4*>(&As[kk * LDA + ty * TM + i]);
          a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[kk * LDA + ty * TM + i];
      }
      {
        const float4 t = *reinterpret_cast<const float4*>(&Bs[kk * LDB + tx * TN]);
        b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if constexpr (Epi::kBias) {
        if (tx == 0 && blockIdx.y == 0) {
#pragma unroll
          for (int i = 0; i < TM; ++i) bsum[i] += a[i];
        }
      }
    }
    __syncthreads();
  }

  const int n = n0 + tx * TN;
  if (gridDim.z > 1) {
    // raw partials: ws[z][M][N] (+ bias partials ws_bias[z][M] after all tile partials)
    float* part = ws + (size_t)blockIdx.z * M * N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty * TM + i;
      if (m < M && n < N)
        *reinterpret_cast<float4*>(&part[(size_t)m * N + n]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    if constexpr (Epi::kBias) {
      if (blockIdx.y == 0 && tx == 0) {
        float* bpart = ws + (size_t)gridDim.z * M * N + (size_t)blockIdx.z * M;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const int m = m0 + ty * TM + i;
          if (m < M) bpart[m] = bsum[i];
        }
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m < M && n < N) epi.store(m, n, acc[i]);
  }
  if constexpr (Epi::kBias) {
    if (blockIdx.y == 0 && tx == 0) {
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m < M) epi.store_bias(m, bsum[i]);
      }
    }
  }
}

template <class Epi>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(Epi epi, const float* __restrict__ ws, int M, int N, int splits) {
  pdl_wait();
  pdl_trigger();
  const size_t total4 = (size_t)M * N / 4;
  const size_t stride = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    float4 s = *reinterpret_cast<const float4*>(ws + i * 4);
    for (int z = 1; z < splits; ++z) {
      const float4 t = *reinterpret_cast<const float4*>(ws + z * stride + i * 4);
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    const int m = (int)((i * 4) / N), n = (int)((i * 4) % N);
    const float v[4] = {s.x, s.y, s.z, s.w};
    epi.store(m, n, v);
  }
  if constexpr (Epi::kBias) {
    const float* bws = ws + (size_t)splits * stride;
    for (size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x; m < (size_t)M; m += (size_t)gridDim.x * blockDim.x) {
      float s = 0.f;
      for (int z = 0; z < splits; ++z) s += bws[(size_t)z * M + m];
      epi.store_bias((int)m, s);
    }
  }
}

}  // namespace clslam
