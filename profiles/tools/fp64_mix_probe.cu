// fp64_mix_probe.cu -- how do DFMA chains fare next to a DMMA stream on the same SM sub-partition?
// One CTA per SM, 8 warps: warps 0-3 (one per sub-partition) stream DMMA.8x8x4 with 4 accumulator
// chains; warps 4-7 (same sub-partitions) run CH independent DFMA chains.  Prints cycles per DFMA
// (per warp) alone and next to the DMMA stream, and what the DMMA stream loses.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_mix_probe fp64_mix_probe.cu && ./fp64_mix_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH, int DC, int OWN>
__global__ void probe(double* out, long long* cyc, int dmma_iters, int dfma_iters, int dmma_on, int gap) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double a = 1.0 + lane * 1e-9, b = 1.0 - lane * 1e-9;
  double s = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < 4) {
    if (dmma_on) {
      double c[DC][2] = {};
      double own = 0.0;
      for (int it = 0; it < dmma_iters * 4 / DC; ++it) {
#pragma unroll
        for (int j = 0; j < DC; ++j) {
          dmma(c[j][0], c[j][1], a, b);
          if (OWN) own = fma(a, own, b);
        }
      }
      for (int j = 0; j < DC; ++j) s += c[j][0] + c[j][1];
      s += own;
    }
  } else {
    double c[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) c[j] = 0.0;
    for (int it = 0; it < dfma_iters; ++it) {
#pragma unroll
      for (int j = 0; j < CH; ++j) c[j] = fma(a, c[j], b);
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) s += c[j];
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (lane == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0;
}

template <int CH, int DC, int OWN>
void run() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 256 * 8);
  cudaMalloc(&cyc, 148 * 8 * 8);
  const int dfma_iters = 20000 / CH;
  long long h[8];
  double alone = 0;
  for (int on = 0; on < 2; ++on) {
    const int dmma_iters = on ? 40000 : 0;   // long enough to cover the DFMA warps
    for (int rep = 0; rep < 2; ++rep) probe<CH, DC, OWN><<<148, 256>>>(out, cyc, dmma_iters, dfma_iters, on, 0);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    const double per = (double)h[4] / (dfma_iters * CH);
    if (!on) alone = per;
    printf("[DMMA chains %d, own DFMA %d] DFMA chains %d %s: %.2f cycles per DFMA per warp%s\n", DC, OWN, CH, on ? "next to a DMMA stream" : "alone", per,
           on ? "" : "");
    if (on) printf("    slowdown %.2fx\n", per / alone);
  }
  // what the DMMA stream loses: DMMA warps timed over a fixed count with the DFMA warps running throughout
  for (int on = 0; on < 2; ++on) {
    const int dmma_iters = 2000;
    for (int rep = 0; rep < 2; ++rep) probe<CH, DC, OWN><<<148, 256>>>(out, cyc, dmma_iters, on ? 400000 / CH : 0, 1, 0);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("    DMMA stream %s: %.2f cycles per DMMA\n", on ? "next to the DFMA warp" : "alone", (double)h[0] / (dmma_iters * 4));
  }
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<1, 4, 0>();
  run<4, 4, 0>();
  run<1, 1, 0>();
  run<4, 1, 0>();
  run<1, 2, 0>();
  run<4, 2, 0>();
  run<1, 7, 0>();
  run<4, 7, 0>();
  run<1, 4, 1>();
  run<4, 4, 1>();
  return 0;
}
