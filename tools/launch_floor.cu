// Where does the ~11 us floor of a conv3x3_tma_kernel launch come from?  Stand-alone measurement (no library code): kernels of
// the same launch shape (148 / 60 / 20 CTAs x 576 threads, ~200 KB dynamic shared memory) with the fixed costs added one at a
// time, timed as a chain of 64 dependent launches inside a CUDA graph (the product's situation) -> microseconds per link.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/launch_floor tools/launch_floor.cu && /tmp/launch_floor
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// LEVEL 0: empty.  1: touches its dynamic shared memory.  2: + 41 mbarrier inits, fence, __syncthreads.  3: + TMEM alloc / dealloc of
// `cols` columns.  4: + one 32 KB bulk copy global -> shared and an "epilogue" (every thread stores 16 bytes).  PDL: griddepcontrol.
template <int LEVEL, bool PDL>
__global__ void __launch_bounds__(576, 1) floor_kernel(const float* __restrict__ src, float* __restrict__ dst, int cols) {
  extern __shared__ uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 512);
  uint8_t* tile = smem + 1024;
  const int warp = threadIdx.x >> 5;
  if (LEVEL >= 1 && threadIdx.x == 0) smem[600] = 1;
  if (LEVEL >= 2) {
    if (threadIdx.x == 0) {
      for (int i = 0; i < 41; ++i) mbar_init(&bars[i], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  if (LEVEL >= 3 && warp == 2) tmem_alloc(slot, (uint32_t)cols);
  if (LEVEL >= 2) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (PDL) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  if (LEVEL >= 4) {
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(&bars[0], 32768);
      bulk_load(tile, src + (size_t)blockIdx.x * 8192, 32768, &bars[0]);
    }
    mbar_wait(&bars[0], 0);
    const float4 v = *reinterpret_cast<const float4*>(tile + threadIdx.x * 16);
    *reinterpret_cast<float4*>(dst + ((size_t)blockIdx.x * 576 + threadIdx.x) * 4) = v;
  }
  if (LEVEL >= 3) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) tmem_dealloc(*slot, (uint32_t)cols);
  }
}

template <int LEVEL, bool PDL>
static float chain_us(int grid, int block, size_t smem, int cols, const float* src, float* dst, cudaStream_t st, int links = 64) {
  auto kern = floor_kernel<LEVEL, PDL>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < links; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = PDL ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, src, dst, cols);
  }
  cudaStreamEndCapture(st, &graph);
  cudaGraphInstantiate(&exec, graph, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(exec, st);
  cudaStreamSynchronize(st);
  float best = 1e9f;
  for (int rep = 0; rep < 10; ++rep) {
    cudaEventRecord(e0, st);
    cudaGraphLaunch(exec, st);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("  (error: %s)\n", cudaGetErrorString(e));
  return best * 1000.f / links;
}

int main() {
  cudaStream_t st;
  cudaStreamCreate(&st);
  float *src, *dst;
  cudaMalloc(&src, 148 * 32768 + 65536);
  cudaMalloc(&dst, 148 * 576 * 16 + 65536);
  cudaMemset(src, 0, 148 * 32768);
  const size_t big = 200 * 1024;
  printf("us per link, chain of 64 dependent launches in one CUDA graph (best of 10 replays)\n");
  printf("%-64s %8s %8s %8s\n", "variant", "148 CTA", "60 CTA", "20 CTA");
  struct Row { const char* name; float v[3]; };
  const int grids[3] = {148, 60, 20};
  auto line = [&](const char* name, auto fn) {
    printf("%-64s", name);
    for (int g = 0; g < 3; ++g) printf(" %8.2f", fn(grids[g]));
    printf("\n");
  };
  line("L0 empty, 576 thr, 0 smem", [&](int g) { return chain_us<0, false>(g, 576, 0, 0, src, dst, st); });
  line("L0 empty, 192 thr, 0 smem", [&](int g) { return chain_us<0, false>(g, 192, 0, 0, src, dst, st); });
  line("L1 576 thr, 200 KB dyn smem", [&](int g) { return chain_us<1, false>(g, 576, big, 0, src, dst, st); });
  line("L1 576 thr, 100 KB dyn smem", [&](int g) { return chain_us<1, false>(g, 576, big / 2, 0, src, dst, st); });
  line("L2 + 41 mbarrier inits + __syncthreads", [&](int g) { return chain_us<2, false>(g, 576, big, 0, src, dst, st); });
  line("L3 + TMEM alloc/dealloc 512 cols", [&](int g) { return chain_us<3, false>(g, 576, big, 512, src, dst, st); });
  line("L3 + TMEM alloc/dealloc 128 cols", [&](int g) { return chain_us<3, false>(g, 576, big, 128, src, dst, st); });
  line("L4 + 32 KB bulk load + 16 B / thread store (512 cols)", [&](int g) { return chain_us<4, false>(g, 576, big, 512, src, dst, st); });
  line("L4 with programmatic dependent launch (512 cols)", [&](int g) { return chain_us<4, true>(g, 576, big, 512, src, dst, st); });
  line("L4 with programmatic dependent launch (128 cols, 100 KB)", [&](int g) { return chain_us<4, true>(g, 576, big / 2, 128, src, dst, st); });
  line("L4 no PDL (128 cols, 100 KB)", [&](int g) { return chain_us<4, false>(g, 576, big / 2, 128, src, dst, st); });
  line("L0 empty with programmatic dependent launch", [&](int g) { return chain_us<0, true>(g, 576, 0, 0, src, dst, st); });
  return 0;
}
