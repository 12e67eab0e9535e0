// predict.cu -- the callers either side of the regression:
//   * kmpc_generate_snapshots: the snapshot generator of data_generate.py:17-74 / 82-152 -- n_traj
//     trajectories x n_step vectorised RK4 steps, written trajectory-major (X, Y, U as the
//     regression reads them) -- so the EDMD stage never leaves the GPU;
//   * kmpc_open_loop_predict: the open-loop multi-step predictor the reference judges the model
//     with (duffing.py:290-343): re-encode from the true state every `reset_every` steps,
//     z+ = A z + B u in between, read-out C z, RMSE of one read-out row.
// Both are HBM-bound streaming kernels: the generator writes 40 B per snapshot (x, y, u) and reads
// 8 B (u0); the predictor reads nz + 1 + n doubles and writes nz + n doubles per step.
#include "common.cuh"
#include "percase.cuh"

namespace kmpc {

constexpr int kGenThreads = 128;   // trajectories per CTA
constexpr int kGenChunk = 8;       // steps staged in shared memory between coalesced write bursts

// One thread integrates one trajectory.  Every kGenChunk steps the CTA flushes its tile
// [trajectory][step][x1, x2 | y1, y2 | u] to the trajectory-major outputs with full-line writes.
__global__ void __launch_bounds__(kGenThreads)
generate_snapshots_kernel(const double* __restrict__ x0, const double* __restrict__ u0,
                          const double* __restrict__ params, int plant_kind, int rk4_variant, double h,
                          int64_t n_traj, int n_step, double* __restrict__ X, double* __restrict__ Y,
                          double* __restrict__ U) {
  __shared__ double sx[kGenThreads][2 * kGenChunk + 1];   // +1: conflict-free column walks
  __shared__ double sy[kGenThreads][2 * kGenChunk + 1];
  __shared__ double su[kGenThreads][kGenChunk + 1];
  const int tid = threadIdx.x;
  const int64_t traj0 = (int64_t)blockIdx.x * kGenThreads;
  const int64_t traj = traj0 + tid;
  const bool valid = traj < n_traj;
  double p[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) p[k] = params[k];
  double x1 = valid ? x0[2 * traj] : 0.0, x2 = valid ? x0[2 * traj + 1] : 0.0;
  const int rows = (int)((n_traj - traj0 < kGenThreads) ? (n_traj - traj0) : kGenThreads);
  for (int j0 = 0; j0 < n_step; j0 += kGenChunk) {
    const int nj = (n_step - j0 < kGenChunk) ? (n_step - j0) : kGenChunk;
    for (int j = 0; j < nj; ++j) {
      const double u = valid ? u0[(int64_t)(j0 + j) * n_traj + traj] : 0.0;   // coalesced over trajectories
      double o1, o2;
      plant_step_dev(plant_kind, rk4_variant, h, p, x1, x2, u, o1, o2);
      sx[tid][2 * j] = x1;
      sx[tid][2 * j + 1] = x2;
      sy[tid][2 * j] = o1;
      sy[tid][2 * j + 1] = o2;
      su[tid][j] = u;
      x1 = o1;
      x2 = o2;
    }
    __syncthreads();
    // trajectory r of the tile owns the contiguous run [(traj0 + r) * n_step + j0, + nj) of snapshots
    for (int e = tid; e < rows * 2 * nj; e += kGenThreads) {
      const int r = e / (2 * nj), c = e - r * (2 * nj);
      const int64_t base = ((traj0 + r) * n_step + j0) * 2 + c;
      X[base] = sx[r][c];
      Y[base] = sy[r][c];
    }
    for (int e = tid; e < rows * nj; e += kGenThreads) {
      const int r = e / nj, c = e - r * nj;
      U[(traj0 + r) * n_step + j0 + c] = su[r][c];
    }
    __syncthreads();
  }
}

// One group of 16 lanes per segment of `reset_every` steps; lane i < nz owns component i of the
// lifted state (row i of A in registers), the state vector is exchanged with shuffles.
constexpr int kPredLanes = 16;
__global__ void __launch_bounds__(128)
open_loop_predict_kernel(const double* __restrict__ psi, const double* __restrict__ u,
                         const double* __restrict__ A, const double* __restrict__ B,
                         const double* __restrict__ C, int nz, int n, int64_t n_seq, int T,
                         int64_t seq_stride, int reset_every, double* __restrict__ decoder_X,
                         double* __restrict__ test_Y) {
  const int lane = threadIdx.x & (kPredLanes - 1);
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kPredLanes;
  const int segs = (T + reset_every - 1) / reset_every;
  const bool live = g < n_seq * segs;
  const int64_t seq = live ? g / segs : 0;
  const int seg = live ? (int)(g - seq * segs) : 0;
  const unsigned mask = 0xffffu << (threadIdx.x & 16);
  double Ar[KMPC_MAX_NZ], Bl = 0.0, Cc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int j = 0; j < KMPC_MAX_NZ; ++j) Ar[j] = (lane < nz && j < nz) ? A[lane * nz + j] : 0.0;
  if (lane < nz) {
    Bl = B[lane];
    for (int r = 0; r < n; ++r) Cc[r] = C[r * nz + lane];   // column `lane` of C
  }
  const int t0 = seg * reset_every;
  const int t1 = (t0 + reset_every < T) ? t0 + reset_every : T;
  const int64_t base = seq * seq_stride;
  // re-encode from the true state (duffing.py:303-305: phix = net.Encoder(inputs_x[:, i]))
  double z = (live && lane < nz) ? psi[(base + t0) * nz + lane] : 0.0;
  for (int t = t0; t < t1; ++t) {
    const double ut = live ? u[base + t] : 0.0;
    if (live && lane < nz) decoder_X[(seq * T + t) * nz + lane] = z;
    // read-out C z: sum over the lanes, in lane order
    for (int r = 0; r < n; ++r) {
      double part = Cc[r] * z, acc = 0.0;
      for (int j = 0; j < nz; ++j) acc += __shfl_sync(mask, part, j, kPredLanes);
      if (live && lane == 0) test_Y[(seq * T + t) * n + r] = acc;
    }
    // z+ = A z + B u (duffing.py:334)
    double zn = 0.0;
#pragma unroll
    for (int j = 0; j < KMPC_MAX_NZ; ++j) {
      const double zj = __shfl_sync(mask, z, j, kPredLanes);
      if (j < nz) zn = fma(Ar[j], zj, zn);
    }
    z = zn + Bl * ut;
  }
}

// RMSE of one read-out row as the reference defines it (duffing.py:341):
// || (test_Y[row] - X[row, :T]) / T ||_2, one warp per sequence, fixed summation order.
__global__ void predict_rmse_kernel(const double* __restrict__ test_Y, const double* __restrict__ x,
                                    int n, int64_t n_seq, int T, int64_t seq_stride, int row,
                                    double* __restrict__ rmse) {
  const int lane = threadIdx.x & 31;
  const int64_t seq = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (seq >= n_seq) return;
  double acc = 0.0;
  for (int t = lane; t < T; t += 32) {
    const double d = (test_Y[(seq * T + t) * n + row] - x[(seq * seq_stride + t) * n + row]) / (double)T;
    acc = fma(d, d, acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) rmse[seq] = sqrt(acc);
}

// Training-loss window sums (duffing.py:179-235).  Window w starts at snapshot k = k0 + w * stride; one warp
// per window, fixed summation order (lane-strided partials, xor-shuffle tree).  out[w] = {rec, lin, pred}:
//   rec  = || Decoder(psi_k) - x_k ||^2                      (xdec[w][0])
//   lin  = sum_{p=1..H} || zpred[w][p] - psi_{k+p} ||^2      (zpred = the linear rollout of the window)
//   pred = sum_{p=1..H} || x_{k+p} - Decoder(zpred[w][p]) ||^2
__global__ void window_losses_kernel(const double* __restrict__ psi, const double* __restrict__ x,
                                     const double* __restrict__ zpred, const double* __restrict__ xdec, int nz,
                                     int n, int64_t W, int T, int64_t k0, int64_t stride,
                                     double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= W) return;
  const int64_t k = k0 + w * stride;
  double rec = 0.0, lin = 0.0, pred = 0.0;
  for (int e = lane; e < n; e += 32) {
    const double d = xdec[(w * T) * n + e] - x[k * n + e];
    rec = fma(d, d, rec);
  }
  for (int e = lane; e < (T - 1) * nz; e += 32) {
    const int p = 1 + e / nz, c = e - (p - 1) * nz;
    const double d = zpred[(w * T + p) * nz + c] - psi[(k + p) * nz + c];
    lin = fma(d, d, lin);
  }
  for (int e = lane; e < (T - 1) * n; e += 32) {
    const int p = 1 + e / n, c = e - (p - 1) * n;
    const double d = x[(k + p) * n + c] - xdec[(w * T + p) * n + c];
    pred = fma(d, d, pred);
  }
  for (int o = 16; o > 0; o >>= 1) {
    rec += __shfl_xor_sync(0xffffffffu, rec, o);
    lin += __shfl_xor_sync(0xffffffffu, lin, o);
    pred += __shfl_xor_sync(0xffffffffu, pred, o);
  }
  if (lane == 0) {
    out[w * 3 + 0] = rec;
    out[w * 3 + 1] = lin;
    out[w * 3 + 2] = pred;
  }
}

}  // namespace kmpc

using namespace kmpc;

extern "C" {

int kmpc_window_losses(const double* psi, const double* x, const double* zpred, const double* xdec, int nz, int n,
                       int64_t W, int T, int64_t k0, int64_t stride, double* out, void* stream) {
  if (nz < 1 || nz > KMPC_MAX_NZ || n < 1 || n > 4 || W < 0 || T < 1 || k0 < 0 || stride < 1) return KMPC_ERR_ARG;
  if (W == 0) return KMPC_OK;
  if (!psi || !x || !zpred || !xdec || !out) return KMPC_ERR_ARG;
  const int64_t blocks = (W * 32 + 127) / 128;
  if (blocks > 0x7fffffff) return KMPC_ERR_ARG;
  window_losses_kernel<<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(psi, x, zpred, xdec, nz, n, W, T, k0, stride, out);
  KMPC_AFTER_LAUNCH();
  return KMPC_OK;
}

int kmpc_generate_snapshots(const double* x0, const double* u0, const double* params, int plant_kind,
                            int rk4_variant, double h, int64_t n_traj, int n_step, double* X, double* Y,
                            double* U, void* stream) {
  if (n_traj < 0 || n_step < 0) return KMPC_ERR_ARG;
  if (plant_kind != KMPC_PLANT_POLY2 && plant_kind != KMPC_PLANT_TANK) return KMPC_ERR_ARG;
  if (n_traj == 0 || n_step == 0) return KMPC_OK;
  if (!x0 || !u0 || !params || !X || !Y || !U) return KMPC_ERR_ARG;
  const int64_t blocks = (n_traj + kGenThreads - 1) / kGenThreads;
  if (blocks > 0x7fffffff) return KMPC_ERR_ARG;
  generate_snapshots_kernel<<<(unsigned)blocks, kGenThreads, 0, as_stream(stream)>>>(
      x0, u0, params, plant_kind, rk4_variant, h, n_traj, n_step, X, Y, U);
  KMPC_AFTER_LAUNCH();
  return KMPC_OK;
}

int kmpc_open_loop_predict(const double* psi, const double* x, const double* u, const double* A,
                           const double* B, const double* C, int nz, int n, int64_t n_seq, int T,
                           int64_t seq_stride, int reset_every, int rmse_row, double* decoder_X,
                           double* test_Y, double* rmse, void* stream) {
  if (nz < 1 || nz > KMPC_MAX_NZ || n < 1 || n > 4 || n_seq < 0 || T < 0 || reset_every < 1) return KMPC_ERR_ARG;
  if (seq_stride < 1) return KMPC_ERR_ARG;   // windows may overlap (training-loss windows: stride 1)
  if (n_seq == 0 || T == 0) return KMPC_OK;
  if (!psi || !u || !A || !B || !C || !decoder_X || !test_Y) return KMPC_ERR_ARG;
  if (rmse && (!x || rmse_row < 0 || rmse_row >= n)) return KMPC_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  const int segs = (T + reset_every - 1) / reset_every;
  const int64_t groups = n_seq * segs;
  const int64_t blocks = (groups * kPredLanes + 127) / 128;
  if (blocks > 0x7fffffff) return KMPC_ERR_ARG;
  open_loop_predict_kernel<<<(unsigned)blocks, 128, 0, st>>>(psi, u, A, B, C, nz, n, n_seq, T, seq_stride,
                                                             reset_every, decoder_X, test_Y);
  KMPC_AFTER_LAUNCH();
  if (rmse) {
    const int64_t rb = (n_seq * 32 + 127) / 128;
    predict_rmse_kernel<<<(unsigned)rb, 128, 0, st>>>(test_Y, x, n, n_seq, T, seq_stride, rmse_row, rmse);
    KMPC_AFTER_LAUNCH();
  }
  return KMPC_OK;
}

}  // extern "C"
