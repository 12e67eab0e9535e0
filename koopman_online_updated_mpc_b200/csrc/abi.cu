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
}  // namespace kmpc

extern "C" {

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
