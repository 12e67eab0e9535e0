// fp64_probe2.cu -- DMMA throughput with the operand pattern of the encoder loop: per k-step two
// A fragments x four B fragments -> 8 DMMAs on 8 accumulators; operands (a) held in registers,
// (b) loaded from shared memory each k-step like the real kernel.  8 warps per SM, 1 CTA per SM.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(double* out, long long* cyc, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 36 * 104 + 100 * 104; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
  const double* ap = sm + tig * 36 + 16 * (warp & 1) + gid;
  const double* bp = sm + 36 * 104 + tig * 100 + (warp >> 1) * 8 + gid;
  double c[2][4][2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[m][j][0] = c[m][j][1] = 0.0;
  double ra0 = 1.0 + lane * 1e-9, ra1 = 1.0 - lane * 1e-9, rb[4] = {1.1, 1.2, 1.3, 1.4};
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 2
    for (int k0 = 0; k0 < 100; k0 += 4) {
      double a0, a1, b[4];
      if (MODE == 0) {
        a0 = ra0; a1 = ra1;
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = rb[j];
      } else {
        a0 = ap[k0 * 36];
        a1 = ap[k0 * 36 + 8];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = (MODE == 2 && j == 3 && (warp >> 1) != 0) ? 0.0 : bp[k0 * 100 + 32 * j > 10000 ? 0 : k0 * 100 + 32 * j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (MODE != 2 || j < 3 || (warp >> 1) == 0) {
          dmma(c[0][j][0], c[0][j][1], a0, b[j]);
          dmma(c[1][j][0], c[1][j][1], a1, b[j]);
        }
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[m][j][0] + c[m][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int sms, double* out, long long* cyc) {
  const int iters = 400, smem = (36 * 104 + 100 * 104) * 8;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<MODE><<<sms, 256, smem>>>(out, cyc, iters);
  probe<MODE><<<sms, 256, smem>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double n_per_smsp = 2.0 * iters * 25 * 8;   // 2 warps per SMSP x 8 DMMA slots per k-step
  printf("%s: %.1f cycles per DMMA slot per SMSP (%lld cycles)\n", name, h / n_per_smsp, h);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * sms * 256);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  run<0>("operands in registers, 8 acc", sms, out, cyc);
  run<1>("operands from shared memory (encoder pattern, 4 n-tiles per warp)", sms, out, cyc);
  run<2>("encoder pattern, 13 n-tiles: warps 0,1 four tiles, others three (slots counted as 4)", sms, out, cyc);
  return 0;
}
