// fp64_peak.cu -- measures the fp64 roofline denominators MEASURED_PEAKS.json does not carry:
// DFMA (CUDA-core) and DMMA (mma.sync.m8n8k4.f64 tensor path) throughput on this GPU.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
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

__global__ void dmma_kernel(double* out, int iters) {
  double c[4][2], a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c[0][0] + c[1][1] + c[2][0] + c[3][1];
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 512);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000, blocks = sms * 8, threads = 512;
  float best_f = 1e30f, best_m = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best_f) best_f = ms;
    cudaEventRecord(e0);
    dmma_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best_m) best_m = ms;
  }
  const double fma_flops = 2.0 * 32.0 * iters * (double)blocks * threads;
  const double mma_flops = 2.0 * 256.0 * 4.0 * iters * (double)blocks * (threads / 32);
  printf("{\"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"dfma_ms\": %.3f, \"dmma_ms\": %.3f}\n", sms,
         fma_flops / best_f * 1e-9, mma_flops / best_m * 1e-9, best_f, best_m);
  return 0;
}
