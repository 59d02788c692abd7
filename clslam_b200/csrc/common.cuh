// Shared helpers for the clslam_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "clslam_b200.h"

#define CLSLAM_HD __host__ __device__ __forceinline__

namespace clslam {

void set_error(const char* fmt, ...);
int finish_launch(const char* what);   // cudaGetLastError -> error string, counts the launch

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch: every kernel is launched with the programmatic-stream-serialization attribute and
// starts with griddepcontrol.wait (all memory of the preceding kernel is visible after it) + launch_dependents, so the
// launch latency and the prologue (barrier init, TMEM allocation, tensor-map prefetch) of kernel i + 1 overlap the
// tail of kernel i.  CLSLAM_PDL=1 in the environment turns the attribute on (default: plain stream order; inside the step graph it measured neutral).
bool pdl_enabled();
template <class... KArgs, class... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);      // errors surface through finish_launch (cudaGetLastError)
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

CLSLAM_HD float4 ld4(const float* p) {
#ifdef __CUDA_ARCH__
  return __ldg(reinterpret_cast<const float4*>(p));
#else
  return *reinterpret_cast<const float4*>(p);
#endif
}
CLSLAM_HD float ld1(const float* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}
CLSLAM_HD float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
CLSLAM_HD void f4add(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// reflection padding index (pad < size): -1 -> 1, size -> size-2   (nn.ReflectionPad2d)
CLSLAM_HD int reflect_idx(int i, int size) {
  if (i < 0) i = -i;
  if (i >= size) i = 2 * size - 2 - i;
  return i;
}

CLSLAM_HD float act_apply(float v, int act) {
  switch (act) {
    case CLSLAM_ACT_RELU: return v > 0.f ? v : 0.f;
    case CLSLAM_ACT_ELU: return v > 0.f ? v : expm1f(v);          // nn.ELU(alpha=1)
    case CLSLAM_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}
// derivative of the activation expressed through its OUTPUT y (what the forward pass stored)
CLSLAM_HD float act_grad_from_output(float y, int act) {
  switch (act) {
    case CLSLAM_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case CLSLAM_ACT_ELU: return y > 0.f ? 1.f : (y + 1.f);        // elu_backward(is_result=True)
    case CLSLAM_ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace clslam
