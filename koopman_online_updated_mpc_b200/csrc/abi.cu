// abi.cu -- error strings, version, launch counter of libkmpc.so.
#include <string.h>

#include "common.cuh"

namespace kmpc {
std::atomic<int64_t> g_launches{0};
static thread_local char g_err[512] = "";
int record_cuda_error(cudaError_t e, const char* what, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  cudaGetLastError();  // clear the (non-sticky) last-error slot so later launches are not blamed
  return KMPC_ERR_CUDA;
}

// ---- fp64 peak probes (roofline denominators; same kernels as profiles/tools/fp64_peak.cu) ------
__global__ void peak_dfma_kernel(double* out, int iters) {
  double a[8], x = 1.0000001, y = 0.9999999;
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void peak_dmma_kernel(double* out, int iters) {
  double c[4][2], a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c[0][0] + c[1][1] + c[2][0] + c[3][1];
}
}  // namespace kmpc

extern "C" {

int kmpc_measure_fp64_peak(double* dmma_tflops, double* dfma_tflops, void* stream) {
  using namespace kmpc;
  if (!dmma_tflops || !dfma_tflops) return KMPC_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  int dev = 0, sms = 0;
  KMPC_CUDA(cudaGetDevice(&dev));
  KMPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int iters = 20000, blocks = sms * 8, threads = 512;
  double* out = nullptr;
  KMPC_CUDA(cudaMalloc(&out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  KMPC_CUDA(cudaEventCreate(&e0));
  KMPC_CUDA(cudaEventCreate(&e1));
  float best_f = 1e30f, best_m = 1e30f;
  int rc = KMPC_OK;
  for (int rep = 0; rep < 3 && rc == KMPC_OK; ++rep) {
    float ms = 0.f;
    cudaEventRecord(e0, st);
    peak_dfma_kernel<<<blocks, threads, 0, st>>>(out, iters);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) rc = KMPC_ERR_CUDA;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best_f) best_f = ms;
    cudaEventRecord(e0, st);
    peak_dmma_kernel<<<blocks, threads, 0, st>>>(out, iters);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) rc = KMPC_ERR_CUDA;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best_m) best_m = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (rc != KMPC_OK) return record_cuda_error(cudaGetLastError(), "fp64 peak probe", __FILE__, __LINE__);
  *dfma_tflops = 2.0 * 32.0 * iters * (double)blocks * threads / best_f * 1e-9;
  *dmma_tflops = 2.0 * 256.0 * 4.0 * iters * (double)blocks * (threads / 32) / best_m * 1e-9;
  return KMPC_OK;
}

const char* kmpc_strerror(int code) {
  switch (code) {
    case KMPC_OK: return "ok";
    case KMPC_ERR_ARG: return "invalid argument";
    case KMPC_ERR_CUDA: return "CUDA error (see kmpc_last_cuda_error)";
    case KMPC_ERR_UNSUPPORTED: return "unsupported configuration";
    case KMPC_ERR_ALLOC: return "allocation failed";
    default: return "unknown error";
  }
}
const char* kmpc_last_cuda_error(void) { return kmpc::g_err; }
int kmpc_version(void) { return KMPC_VERSION; }
int64_t kmpc_launch_count(void) { return kmpc::g_launches.load(); }

}  // extern "C"
