// stages.cu -- stage-level entry points (one reference call site each): RLS update, MPC first
// move, plant step, RBF lift.  Warp-per-scenario kernels built from percase.cuh.
#include "common.cuh"
#include "percase.cuh"

namespace kmpc {

// ------------------------------------------------------------------------------ RLS ----------
// grid: ceil(S / warps_per_block) blocks of kWarpsPerBlock warps; smem: one RlsWs per warp.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rls_update_kernel(double* __restrict__ KA, double* __restrict__ P, double* __restrict__ barX,
                  double* __restrict__ barQ, const double* __restrict__ z,
                  const double* __restrict__ u, const double* __restrict__ y,
                  const double* __restrict__ xc, double* __restrict__ A, double* __restrict__ B,
                  double* __restrict__ C, int64_t S, int nz, int n, double lam, int flags) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (s >= S) return;
  const int nv = nz + 1;
  RlsWs ws = rls_ws_carve(smem + (size_t)warp * rls_ws_doubles(nz, n), nz, n);
  for (int e = lane; e < nz * nv; e += 32) ws.KA[e] = KA[s * nz * nv + e];
  for (int e = lane; e < nv * nv; e += 32) ws.P[e] = P[s * nv * nv + e];
  if (flags & KMPC_RLS_UPDATE_C) {
    for (int e = lane; e < n * nz; e += 32) ws.barX[e] = barX[s * n * nz + e];
    for (int e = lane; e < nz * nz; e += 32) ws.barQ[e] = barQ[s * nz * nz + e];
    for (int e = lane; e < n; e += 32) ws.xc[e] = xc[s * n + e];
  }
  for (int e = lane; e < nz; e += 32) {
    ws.v[e] = z[s * nz + e];
    ws.y[e] = y[s * nz + e];
  }
  if (lane == 0) ws.v[nz] = u[s];
  __syncwarp();
  rls_update_warp<32>(ws, nz, n, lam, flags);
  for (int e = lane; e < nz * nz; e += 32) A[s * nz * nz + e] = ws.oA[e];
  for (int e = lane; e < nz; e += 32) B[s * nz + e] = ws.oB[e];
  if (flags & KMPC_RLS_UPDATE_C)
    for (int e = lane; e < n * nz; e += 32) C[s * n * nz + e] = ws.oC[e];
  for (int e = lane; e < nz * nv; e += 32) KA[s * nz * nv + e] = ws.KA[e];
  for (int e = lane; e < nv * nv; e += 32) P[s * nv * nv + e] = ws.P[e];
  if (flags & KMPC_RLS_UPDATE_C) {
    for (int e = lane; e < n * nz; e += 32) barX[s * n * nz + e] = ws.barX[e];
    for (int e = lane; e < nz * nz; e += 32) barQ[s * nz * nz + e] = ws.barQ[e];
  }
}

// ------------------------------------------------------------------------------ QP -----------
// G lanes per scenario.
template <int G>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
qp_first_move_kernel(const double* __restrict__ A, const double* __restrict__ B,
                     const double* __restrict__ Cy, const double* __restrict__ z0,
                     const double* __restrict__ r, const double* __restrict__ lb,
                     const double* __restrict__ ub, const double* __restrict__ PN, double q,
                     double rw, int N, int ny, int nz, int64_t S, int flags,
                     double* __restrict__ u0, double* __restrict__ Ufull, int* __restrict__ status,
                     int max_iter, double tol) {
  extern __shared__ double smem[];
  const int group = threadIdx.x / G, lane = threadIdx.x & (G - 1);
  int64_t s = (int64_t)blockIdx.x * (blockDim.x / G) + group;
  const bool valid = s < S;
  if (!valid) s = S - 1;
  const bool identity = flags & KMPC_QP_CY_IDENTITY;
  const bool shared_model = flags & KMPC_QP_SHARED_MODEL;
  const bool r_full = flags & KMPC_QP_R_FULL;
  QpWs ws = qp_ws_carve(smem + (size_t)group * qp_ws_doubles(nz, ny, N, identity), nz, ny, N, identity);
  const int64_t sm = shared_model ? 0 : s;
  for (int e = lane; e < nz * nz; e += G) ws.A[e] = A[sm * nz * nz + e];
  for (int e = lane; e < nz; e += G) {
    ws.B[e] = B[sm * nz + e];
    ws.z0[e] = z0[s * nz + e];
  }
  if (!identity)
    for (int e = lane; e < ny * nz; e += G) ws.Cy[e] = Cy[sm * ny * nz + e];
  for (int e = lane; e < N; e += G) {
    ws.lb[e] = lb[s * N + e];
    ws.ub[e] = ub[s * N + e];
  }
  __syncwarp();
  const double* rs = r_full ? r + s * N * ny : r + s * ny;
  const double* pn = PN ? PN + sm * ny * ny : nullptr;
  qp_build_warp<G>(ws, nz, ny, N, identity, q, rw, rs, r_full ? ny : 0, pn);
  const int st = qp_solve_warp<G>(ws, N, max_iter, tol);
  if (!valid) return;
  if (lane == 0) {
    u0[s] = ws.x[0];
    if (status) status[s] = st;
  }
  if (Ufull)
    for (int e = lane; e < N; e += G) Ufull[s * N + e] = ws.x[e];
}

// ------------------------------------------------------------------------------ plant --------
__global__ void plant_step_kernel(const double* __restrict__ x, const double* __restrict__ u,
                                  const double* __restrict__ params, double* __restrict__ xnext,
                                  int64_t S, int kind, int rk4_variant, double h) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double p[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) p[k] = params[s * 5 + k];
  double o1, o2;
  plant_step_dev(kind, rk4_variant, h, p, x[2 * s], x[2 * s + 1], u[s], o1, o2);
  xnext[2 * s] = o1;
  xnext[2 * s + 1] = o2;
}

// ------------------------------------------------------------------------------ RBF ----------
__global__ void rbf_lift_kernel(const double* __restrict__ x, const double* __restrict__ cx,
                                double* __restrict__ z, int64_t S, int n, int nz, int variant) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= S * nz) return;
  const int64_t s = e / nz;
  const int c = (int)(e - s * nz);
  double xv[4];
  for (int k = 0; k < n; ++k) xv[k] = x[s * n + k];
  z[e] = rbf_thinplate(xv, cx + c * n, n, variant);
}

}  // namespace kmpc

using namespace kmpc;

extern "C" {

int kmpc_rls_update(double* KA, double* P, double* barX, double* barQ, const double* z,
                    const double* u, const double* y, const double* xc, double* A, double* B,
                    double* C, int64_t S, int nz, int n, double lambda, int flags, void* stream) {
  if (nz < 1 || nz > KMPC_MAX_NZ || n < 1 || n > 4 || !(lambda > 0.0) || S < 0) return KMPC_ERR_ARG;
  if (S == 0) return KMPC_OK;  // empty batch: nothing to validate (empty tensors have null data)
  if (!KA || !P || !z || !u || !y || !A || !B) return KMPC_ERR_ARG;
  if ((flags & KMPC_RLS_UPDATE_C) && (!barX || !barQ || !xc || !C)) return KMPC_ERR_ARG;
  const int wpb = warps_that_fit(rls_ws_doubles(nz, n));
  if (wpb < 1) return KMPC_ERR_UNSUPPORTED;
  const int smem = wpb * rls_ws_doubles(nz, n) * (int)sizeof(double);
  KMPC_CUDA(ensure_smem(rls_update_kernel, smem));
  const unsigned grid = (unsigned)((S + wpb - 1) / wpb);
  rls_update_kernel<<<grid, wpb * 32, smem, as_stream(stream)>>>(
      KA, P, barX, barQ, z, u, y, xc, A, B, C, S, nz, n, lambda, flags);
  KMPC_AFTER_LAUNCH();
  return KMPC_OK;
}

int kmpc_qp_first_move(const double* A, const double* B, const double* Cy, const double* z0,
                       const double* r, const double* lb, const double* ub, const double* PN,
                       double q, double rw, int N, int ny, int nz, int64_t S, int flags,
                       double* u0, double* Ufull, int* status, int max_iter, double tol,
                       void* stream) {
  const bool identity = flags & KMPC_QP_CY_IDENTITY;
  if (identity && ny != nz) return KMPC_ERR_ARG;
  if (nz < 1 || nz > KMPC_MAX_NZ || ny < 1 || ny > KMPC_MAX_NZ || N < 1 || N > KMPC_MAX_HORIZON || S < 0)
    return KMPC_ERR_ARG;
  if (S == 0) return KMPC_OK;
  if (!A || !B || !z0 || !r || !lb || !ub || !u0) return KMPC_ERR_ARG;
  if (!identity && !Cy) return KMPC_ERR_ARG;
  if (max_iter <= 0) max_iter = 10 * N + 20;
  if (!(tol > 0.0)) tol = 1e-10;
  // horizon 10 (duffing.py / vanderpol.py): 16 lanes per scenario, two scenarios per warp
  typedef void (*Kern)(const double*, const double*, const double*, const double*, const double*,
                       const double*, const double*, const double*, double, double, int, int, int,
                       int64_t, int, double*, double*, int*, int, double);
  Kern kern = qp_first_move_kernel<32>;
  int g = 32;
  if (N == 10) {
    kern = qp_first_move_kernel<16>;
    g = 16;
  }
  const int ws_bytes = qp_ws_doubles(nz, ny, N, identity) * (int)sizeof(double);
  int spb = (kWarpsPerBlock * 32) / g;
  while (spb > 32 / g && spb * ws_bytes > 200 * 1024) spb >>= 1;
  if (spb * ws_bytes > 200 * 1024) return KMPC_ERR_UNSUPPORTED;
  const int smem = spb * ws_bytes;
  KMPC_CUDA(ensure_smem(kern, smem));
  const unsigned grid = (unsigned)((S + spb - 1) / spb);
  kern<<<grid, spb * g, smem, as_stream(stream)>>>(A, B, Cy, z0, r, lb, ub, PN, q, rw, N, ny, nz, S, flags, u0,
                                                  Ufull, status, max_iter, tol);
  KMPC_AFTER_LAUNCH();
  return KMPC_OK;
}

int kmpc_plant_step(const double* x, const double* u, const double* params, double* xnext,
                    int64_t S, int kind, int rk4_variant, double h, void* stream) {
  if (kind != KMPC_PLANT_POLY2 && kind != KMPC_PLANT_TANK) return KMPC_ERR_ARG;
  if (S < 0) return KMPC_ERR_ARG;
  if (S == 0) return KMPC_OK;
  if (!x || !u || !params || !xnext) return KMPC_ERR_ARG;
  const unsigned grid = (unsigned)((S + 127) / 128);
  plant_step_kernel<<<grid, 128, 0, as_stream(stream)>>>(x, u, params, xnext, S, kind, rk4_variant, h);
  KMPC_AFTER_LAUNCH();
  return KMPC_OK;
}

int kmpc_rbf_lift(const double* x, const double* cx, double* z, int64_t S, int n, int nz,
                  int variant, void* stream) {
  if (S < 0 || n < 1 || n > 4 || nz < 1) return KMPC_ERR_ARG;
  if (variant != KMPC_RBF_PYTHON && variant != KMPC_RBF_MATLAB) return KMPC_ERR_ARG;
  if (S == 0) return KMPC_OK;
  if (!x || !cx || !z) return KMPC_ERR_ARG;
  const int64_t total = S * nz;
  const unsigned grid = (unsigned)((total + 255) / 256);
  KMPC_CUDA(launch_pdl(rbf_lift_kernel, grid, 256u, (size_t)0, as_stream(stream), x, cx, z, S, n, nz, variant));
  KMPC_AFTER_LAUNCH();
  return KMPC_OK;
}

}  // extern "C"
