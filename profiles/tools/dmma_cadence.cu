// dmma_cadence.cu -- how fast can ONE warp issue fp64 tensor-path MMAs (mma.sync.m8n8k4.f64, SASS
// DMMA.8x8x4) on B200, as a function of (a) warps per SM sub-partition and (b) independent
// accumulator chains per warp?  Prints cycles per DMMA per warp and the SM-wide DMMA rate.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_cadence dmma_cadence.cu && ./dmma_cadence
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void probe(double* out, long long* cyc, int iters) {
  const int lane = threadIdx.x & 31;
  double c[CHAINS][2];
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) c[j][0] = c[j][1] = 0.0;
  const double a = 1.0 + lane * 1e-9, b = 1.0 - lane * 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) dmma(c[j][0], c[j][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (lane == 0) cyc[blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)] = t1 - t0;
}

// DFMA chains for comparison: CHAINS independent dependent-FMA chains per thread
template <int CHAINS>
__global__ void probe_dfma(double* out, long long* cyc, int iters) {
  const int lane = threadIdx.x & 31;
  double c[CHAINS];
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) c[j] = 0.0;
  const double a = 1.0 + lane * 1e-9, b = 1.0 - lane * 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) c[j] = fma(a, c[j], b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (lane == 0) cyc[blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)] = t1 - t0;
}

template <int CHAINS>
void run(int warps, bool dfma) {
  double* out;
  long long* cyc;
  const int iters = 2000;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 148 * 32 * 8);
  for (int rep = 0; rep < 2; ++rep) {
    if (dfma) probe_dfma<CHAINS><<<148, warps * 32>>>(out, cyc, iters);
    else probe<CHAINS><<<148, warps * 32>>>(out, cyc, iters);
  }
  cudaDeviceSynchronize();
  long long h[32];
  cudaMemcpy(h, cyc, warps * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
  const double per = (double)mx / (iters * CHAINS);
  printf("%s warps/SM %2d (%.1f per sub-partition) chains %d: %.1f cycles per %s per warp; per sub-partition one every %.1f cycles\n",
         dfma ? "DFMA" : "DMMA", warps, warps / 4.0, CHAINS, per, dfma ? "DFMA" : "DMMA", per / (warps / 4.0 < 1 ? 1 : warps / 4.0));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {1, 4, 8, 12, 16, 32}) {
    run<1>(w, false);
    run<2>(w, false);
    run<4>(w, false);
    run<7>(w, false);
  }
  for (int w : {1, 4, 8, 16}) {
    run<1>(w, true);
    run<4>(w, true);
    run<8>(w, true);
  }
  return 0;
}
