// common.cuh -- host-side helpers shared by the translation units of libkmpc.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>

#include "../../include/kmpc.h"

namespace kmpc {

extern std::atomic<int64_t> g_launches;
int record_cuda_error(cudaError_t e, const char* what, const char* file, int line);

#define KMPC_CUDA(expr)                                                              \
  do {                                                                               \
    cudaError_t kmpc_e_ = (expr);                                                    \
    if (kmpc_e_ != cudaSuccess)                                                      \
      return ::kmpc::record_cuda_error(kmpc_e_, #expr, __FILE__, __LINE__);          \
  } while (0)

// after every <<<>>>: count the launch and surface launch-configuration errors
#define KMPC_AFTER_LAUNCH()                                                          \
  do {                                                                               \
    ::kmpc::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
    cudaError_t kmpc_e_ = cudaGetLastError();                                        \
    if (kmpc_e_ != cudaSuccess)                                                      \
      return ::kmpc::record_cuda_error(kmpc_e_, "kernel launch", __FILE__, __LINE__); \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// opt in to > 48 KB dynamic shared memory once per kernel
template <typename K>
inline cudaError_t ensure_smem(K kernel, int bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

constexpr int kWarpsPerBlock = 4;

// ---- programmatic dependent launch (PDL): the next kernel of the step is launched while the
// current one is still running; it may do work that does not depend on its predecessor (weight
// staging, RLS state loads) and then blocks in pdl_wait() until the predecessor has completed and
// its writes are visible.  Every kernel triggers its successor only AFTER its own wait, so a
// prologue can overlap the immediate predecessor only.  SASS: ACQBULK / PREEXIT.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem,
                              cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// how many warps (each with `doubles_per_warp` of workspace) fit in a block's shared memory
inline int warps_that_fit(int doubles_per_warp) {
  const int budget = 200 * 1024;
  int w = kWarpsPerBlock;
  while (w > 1 && w * doubles_per_warp * (int)sizeof(double) > budget) w >>= 1;
  return (w * doubles_per_warp * (int)sizeof(double) <= budget) ? w : 0;
}

}  // namespace kmpc
