// edmd.cu -- EDMD regression: Gram accumulation over snapshots and the small SPD solves.
//
// Reference: duffing.py:167-177 ([A B] = PHIY pinv([PHIX;U]), C = X pinv(PHIX)) and the Gram form
// Tank_System.m:93-100 (M = (W V') pinv(V V')).  With V full row rank (true for every named
// config, cond(VV') ~ 3e3) pinv(V V') == inv(V V') and the two forms agree.
//
// pack layout (doubles): G = V V' (nv*nv) | Aq = PHIY V' (nz*nv) | XV = X V' (n*nv) | count.
// The rows R = [V; PHIY; X] (nr = nv + nz + n) are multiplied against V' : out[rr][c].
#include "common.cuh"
#include "percase.cuh"

namespace kmpc {

// Fast path: each thread owns snapshots m = base + k*stride and a block of kCB = 3 columns of V;
// `roles` = ceil(nv / 3) groups of 64 threads cover all columns.  All nr x 3 partial sums live
// in registers (fp64).  Threads of different roles re-read the same snapshot rows (L1 hits).
template <int NZ, int NX>
__global__ void __launch_bounds__(64 * ((NZ + 1 + 2) / 3))
gram_kernel(const double* __restrict__ psi, const double* __restrict__ psin,
            const double* __restrict__ u, const double* __restrict__ x, int64_t M,
            double* __restrict__ pack, int seg) {
  constexpr int NV = NZ + 1, NR = NV + NZ + NX, CB = 3, ROLES = (NV + CB - 1) / CB;
  __shared__ double red[ROLES][NR * CB];
  const int role = threadIdx.x / 64, tl = threadIdx.x % 64;
  const int c0 = role * CB;
  double acc[NR][CB];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int c = 0; c < CB; ++c) acc[r][c] = 0.0;
  for (int64_t m = (int64_t)blockIdx.x * 64 + tl; m < M; m += (int64_t)gridDim.x * 64) {
    double row[NR];
    static_assert(NZ % 2 == 0, "vectorised loads need even NZ");
    // seg > 0: trajectory layout -- the lifted rows of a trajectory of `seg` snapshots are its
    // seg + 1 consecutive states, so snapshot m reads rows m + m / seg and the one after it
    const int64_t pr = seg ? m + m / seg : m;
    const double* prow = psi + pr * NZ;
    const double* nrow = seg ? prow + NZ : psin + m * NZ;
#pragma unroll
    for (int k = 0; k < NZ; k += 2) {
      const double2 a = __ldg(reinterpret_cast<const double2*>(prow + k));
      row[k] = a.x;
      row[k + 1] = a.y;
      const double2 b = __ldg(reinterpret_cast<const double2*>(nrow + k));
      row[NV + k] = b.x;
      row[NV + k + 1] = b.y;
    }
    row[NZ] = __ldg(u + m);
#pragma unroll
    for (int k = 0; k < NX; ++k) row[NV + NZ + k] = __ldg(x + m * NX + k);
    double vc[CB];  // this role's columns of V, re-read (L1 hit) to avoid dynamic register indexing
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      const int col = c0 + c;
      vc[c] = col < NZ ? __ldg(prow + col) : (col == NZ ? row[NZ] : 0.0);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int c = 0; c < CB; ++c) acc[r][c] = fma(row[r], vc[c], acc[r][c]);
  }
  // warp reduce, then the two warps of a role combine through shared memory
  const int lane = threadIdx.x & 31, wir = (threadIdx.x >> 5) & 1;
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      double v = acc[r][c];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      acc[r][c] = v;
    }
  if (lane == 0 && wir == 0)
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int c = 0; c < CB; ++c) red[role][r * CB + c] = acc[r][c];
  __syncthreads();
  if (lane == 0 && wir == 1)
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int c = 0; c < CB; ++c) red[role][r * CB + c] += acc[r][c];
  __syncthreads();
  for (int e = threadIdx.x; e < NR * NV; e += blockDim.x) {
    const int r = e / NV, c = e - r * NV;
    atomicAdd(pack + e, red[c / CB][r * CB + (c % CB)]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(pack + NR * NV, (double)M);   // snapshot count (exact)
}

// Generic path (any nz <= KMPC_MAX_NZ): one thread per output element and block-strided
// snapshots; used for dimensions without a specialised instantiation.
__global__ void gram_generic_kernel(const double* __restrict__ psi, const double* __restrict__ psin,
                                    const double* __restrict__ u, const double* __restrict__ x,
                                    int64_t M, int nz, int n, double* __restrict__ pack, int seg) {
  const int nv = nz + 1, nr = nv + nz + n;
  const int e = threadIdx.x;
  if (e >= nr * nv) return;
  const int r = e / nv, c = e - r * nv;
  double acc = 0.0;
  for (int64_t m = blockIdx.x; m < M; m += gridDim.x) {
    const int64_t pr = seg ? m + m / seg : m;
    const double* nrow = seg ? psi + (pr + 1) * nz : psin + m * nz;
    const double vc = (c < nz) ? psi[pr * nz + c] : u[m];
    double vr;
    if (r < nz) vr = psi[pr * nz + r];
    else if (r == nz) vr = u[m];
    else if (r < nv + nz) vr = nrow[r - nv];
    else vr = x[m * n + (r - nv - nz)];
    acc = fma(vr, vc, acc);
  }
  atomicAdd(pack + e, acc);
}

__global__ void gram_count_kernel(double* pack, int idx, double m) { pack[idx] += m; }

// EDMD solve, one warp.  smem: G copy (nv*nv) + rows ((nz+n)*nv) + G2 (nz*nz) + rows2 (n*nz).
__global__ void edmd_solve_kernel(const double* __restrict__ pack, int nz, int n, int c_variant,
                                  double* __restrict__ A, double* __restrict__ B,
                                  double* __restrict__ C, int* __restrict__ status) {
  extern __shared__ double smem[];
  const int nv = nz + 1, lane = threadIdx.x;
  double* G = smem;
  double* R = G + nv * nv;            // (nz + n) x nv : [Aq; XV]
  double* G2 = R + (nz + n) * nv;     // nz x nz
  double* R2 = G2 + nz * nz;          // n x nz
  for (int e = lane; e < nv * nv; e += 32) G[e] = pack[e];
  for (int e = lane; e < (nz + n) * nv; e += 32) R[e] = pack[nv * nv + e];
  for (int e = lane; e < nz * nz; e += 32) G2[e] = pack[(e / nz) * nv + (e % nz)];
  for (int e = lane; e < n * nz; e += 32) R2[e] = pack[nv * nv + nz * nv + (e / nz) * nv + (e % nz)];
  __syncwarp();
  int st = spd_right_solve_warp<32>(G, nv, R, nz + n);   // [A B; Cj *] = [Aq; XV] G^-1
  if (c_variant == KMPC_C_PYTHON) st |= spd_right_solve_warp<32>(G2, nz, R2, n);
  for (int e = lane; e < nz * nv; e += 32) {
    const int i = e / nv, j = e - i * nv;
    if (j < nz) A[i * nz + j] = R[e];
    else B[i] = R[e];
  }
  for (int e = lane; e < n * nz; e += 32) {
    const int i = e / nz, j = e - i * nz;
    C[e] = (c_variant == KMPC_C_PYTHON) ? R2[e] : R[(nz + i) * nv + j];
  }
  if (lane == 0 && status) *status = st;
}

}  // namespace kmpc

using namespace kmpc;

extern "C" {

int64_t kmpc_gram_pack_len(int nz, int n) {
  const int nv = nz + 1;
  return (int64_t)(nv + nz + n) * nv + 1;
}

}  // extern "C"

namespace kmpc {
// seg == 0: psi / psi_next are (M, nz); seg > 0: psi holds M / seg trajectories of seg + 1 lifted
// states each and psi_next is ignored (lift.cu: kmpc_gram_from_trajectories)
int gram_accumulate_impl(const double* psi, const double* psi_next, const double* u, const double* x,
                         int64_t M, int nz, int n, double* pack, int seg, void* stream) {
  if (nz < 1 || nz > KMPC_MAX_NZ || n < 1 || n > 4 || M < 0 || seg < 0) return KMPC_ERR_ARG;
  if (M == 0) return KMPC_OK;
  if (!psi || (!psi_next && !seg) || !u || !x || !pack) return KMPC_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  int dev = 0, sms = 148;
  KMPC_CUDA(cudaGetDevice(&dev));
  KMPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int64_t want = (M + 63) / 64;
  const unsigned grid = (unsigned)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
  if (nz == 8 && n == 2) {
    gram_kernel<8, 2><<<grid, 64 * 3, 0, st>>>(psi, psi_next, u, x, M, pack, seg);
    KMPC_AFTER_LAUNCH();
    return KMPC_OK;
  } else if (nz == 10 && n == 2) {
    gram_kernel<10, 2><<<grid, 64 * 4, 0, st>>>(psi, psi_next, u, x, M, pack, seg);
    KMPC_AFTER_LAUNCH();
    return KMPC_OK;
  } else {
    const int nv = nz + 1, outs = (nv + nz + n) * nv;
    const unsigned g2 = (unsigned)(M < (int64_t)sms * 4 ? M : (int64_t)sms * 4);
    gram_generic_kernel<<<g2, (outs + 31) / 32 * 32, 0, st>>>(psi, psi_next, u, x, M, nz, n, pack, seg);
  }
  KMPC_AFTER_LAUNCH();
  gram_count_kernel<<<1, 1, 0, st>>>(pack, (int)kmpc_gram_pack_len(nz, n) - 1, (double)M);
  KMPC_AFTER_LAUNCH();
  return KMPC_OK;
}
}  // namespace kmpc

extern "C" {

int kmpc_gram_accumulate(const double* psi, const double* psi_next, const double* u,
                         const double* x, int64_t M, int nz, int n, double* pack, void* stream) {
  return gram_accumulate_impl(psi, psi_next, u, x, M, nz, n, pack, 0, stream);
}

int kmpc_edmd_solve(const double* pack, int nz, int n, int c_variant, double* A, double* B,
                    double* C, int* status, void* stream) {
  if (!pack || !A || !B || !C) return KMPC_ERR_ARG;
  if (nz < 1 || nz > KMPC_MAX_NZ || n < 1 || n > 4) return KMPC_ERR_ARG;
  if (c_variant != KMPC_C_PYTHON && c_variant != KMPC_C_JOINT) return KMPC_ERR_ARG;
  const int nv = nz + 1;
  const int smem = (nv * nv + (nz + n) * nv + nz * nz + n * nz) * (int)sizeof(double);
  edmd_solve_kernel<<<1, 32, smem, as_stream(stream)>>>(pack, nz, n, c_variant, A, B, C, status);
  KMPC_AFTER_LAUNCH();
  return KMPC_OK;
}

}  // extern "C"
