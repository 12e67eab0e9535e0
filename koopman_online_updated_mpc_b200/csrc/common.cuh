// common.cuh -- host-side helpers shared by the translation units of libkmpc.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

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

// how many warps (each with `doubles_per_warp` of workspace) fit in a block's shared memory
inline int warps_that_fit(int doubles_per_warp) {
  const int budget = 200 * 1024;
  int w = kWarpsPerBlock;
  while (w > 1 && w * doubles_per_warp * (int)sizeof(double) > budget) w >>= 1;
  return (w * doubles_per_warp * (int)sizeof(double) <= budget) ? w : 0;
}

}  // namespace kmpc
