// fp64_probe.cu -- DMMA (mma.sync.m8n8k4.f64) issue behaviour on one SM configuration:
// 1 CTA of W warps per SM, each warp NACC independent accumulators.  Reports cycles per DMMA per
// SM sub-partition (4 SMSPs, W/4 warps each) and the SM clock seen during the run.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_probe fp64_probe.cu && ./fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void dmma_probe(double* out, long long* cyc, int iters) {
  double c[NACC][2], a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NACC>
__global__ void dfma_probe(double* out, long long* cyc, int iters) {
  double c[NACC], a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename K>
void run(const char* name, K kern, int nacc, int warps, int iters, int sms, double* out, long long* cyc, bool mma) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  kern<<<sms, warps * 32>>>(out, cyc, iters);
  cudaEventRecord(e0);
  kern<<<sms, warps * 32>>>(out, cyc, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double per_smsp = (double)h / ((double)iters * nacc * (warps / 4.0 < 1 ? 1 : warps / 4.0));
  const double flops = (mma ? 512.0 : 64.0) * iters * nacc * warps * sms;
  printf("%s nacc=%d warps=%d: %.1f cycles per instr per SMSP, %.2f TFLOP/s, clock %.0f MHz\n", name, nacc, warps,
         per_smsp, flops / ms * 1e-9, h / ms * 1e-3);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * sms * 1024);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    run("dmma", dmma_probe<1>, 1, warps, iters, sms, out, cyc, true);
    run("dmma", dmma_probe<2>, 2, warps, iters, sms, out, cyc, true);
    run("dmma", dmma_probe<4>, 4, warps, iters, sms, out, cyc, true);
    run("dmma", dmma_probe<8>, 8, warps, iters, sms, out, cyc, true);
  }
  for (int warps : {4, 8, 16, 32}) {
    run("dfma", dfma_probe<1>, 1, warps, iters, sms, out, cyc, false);
    run("dfma", dfma_probe<4>, 4, warps, iters, sms, out, cyc, false);
    run("dfma", dfma_probe<8>, 8, warps, iters, sms, out, cyc, false);
  }
  return 0;
}
