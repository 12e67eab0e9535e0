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

// nz = 8, n = 2 on the fp64 tensor path (mma.sync m8n8k4 f64, SASS DMMA): the Gram is the skinny GEMM
// R V' with K = snapshots.  A warp walks groups of 4 snapshots (one k-step): lane (gid, tig) = (component,
// snapshot in the group) loads ONE double of the lifted state and one of its successor -- a group is 256
// contiguous bytes of each -- and these registers are at once the A fragments of the row tiles
// [z], [z+], [u, x1, x2, 0...] and the B fragments of the column tiles [z], [u, 0...]: six DMMAs per
// group, no shared memory, 12 accumulator doubles per lane for the whole launch.  fp64 accumulation
//  on 24 pipe cycles per snapshot and sub-partition instead of 171 FMAs + 60 loads per
// snapshot on CUDA cores.
__device__ __forceinline__ void gram_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

constexpr int kGramDmmaWarps = 8;

__global__ void __launch_bounds__(kGramDmmaWarps * 32, 2)
gram_dmma_kernel(const double* __restrict__ psi, const double* __restrict__ psin, const double* __restrict__ u,
                 const double* __restrict__ x, int64_t M, double* __restrict__ pack, int seg) {
  constexpr int NZ = 8, NV = 9, NR = 19;
  __shared__ double red[NR * NV];
  for (int e = threadIdx.x; e < NR * NV; e += blockDim.x) red[e] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const int64_t warp = (int64_t)blockIdx.x * kGramDmmaWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kGramDmmaWarps;
  const int64_t groups = (M + 3) >> 2;
  // a warp owns a CONTIGUOUS range of groups (streaming reads; the trajectory index of the lane's snapshot
  // advances incrementally: one division per lane and launch instead of one per group)
  const int64_t per = (groups + nwarps - 1) / nwarps;
  const int64_t g0 = warp * per, g1 = (g0 + per < groups) ? g0 + per : groups;
  double c[3][2][2];
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int n = 0; n < 2; ++n) c[t][n][0] = c[t][n][1] = 0.0;
  int64_t m = 4 * g0 + tig;                 // this lane's snapshot of the group being LOADED
  int64_t traj = seg ? m / seg : 0;
  int off = seg ? (int)(m - traj * seg) : 0;
  int64_t gl = g0;                          // group being loaded
  // fragments of the next group; snapshots beyond M (or beyond the warp's range) contribute zeros
  auto load = [&](double& az, double& an, double& am) {
    az = an = am = 0.0;
    if (gl < g1 && m < M) {
      const int64_t pr = m + traj;          // trajectory layout: seg + 1 consecutive states per trajectory
      az = __ldg(psi + pr * NZ + gid);
      an = __ldg((seg ? psi + (pr + 1) * NZ : psin + m * NZ) + gid);
      if (gid == 0) am = __ldg(u + m);
      else if (gid < 3) am = __ldg(x + m * 2 + (gid - 1));
    }
    ++gl;
    m += 4;
    if (seg) {
      off += 4;
      while (off >= seg) {
        off -= seg;
        ++traj;
      }
    }
  };
  auto mult = [&](double az, double an, double am) {
    const double b1 = (gid == 0) ? am : 0.0;
    gram_dmma(c[0][0][0], c[0][0][1], az, az);
    gram_dmma(c[1][0][0], c[1][0][1], an, az);
    gram_dmma(c[2][0][0], c[2][0][1], am, az);
    gram_dmma(c[0][1][0], c[0][1][1], az, b1);
    gram_dmma(c[1][1][0], c[1][1][1], an, b1);
    gram_dmma(c[2][1][0], c[2][1][1], am, b1);
  };
  // four groups in flight per warp: the loads of groups g + 4 .. g + 7 fly while g .. g + 3 multiply
  double f[4][3], h[4][3];
#pragma unroll
  for (int k = 0; k < 4; ++k) load(f[k][0], f[k][1], f[k][2]);
  for (int64_t g = g0; g < g1; g += 8) {
#pragma unroll
    for (int k = 0; k < 4; ++k) load(h[k][0], h[k][1], h[k][2]);
#pragma unroll
    for (int k = 0; k < 4; ++k) mult(f[k][0], f[k][1], f[k][2]);
#pragma unroll
    for (int k = 0; k < 4; ++k) load(f[k][0], f[k][1], f[k][2]);
#pragma unroll
    for (int k = 0; k < 4; ++k) mult(h[k][0], h[k][1], h[k][2]);
  }
  // accumulator (tile t, column tile n): lane holds rows gid, columns 2 tig, 2 tig + 1 -> pack positions
  //   tile 0 = z rows -> G rows 0..7; tile 1 = z+ rows -> Aq rows; tile 2 = [u, x1, x2] -> G row 8, XV rows
  //   column tile 0 = V columns 0..7, column tile 1 = column 8 (u) in its column 0
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = n ? 8 + 2 * tig + h : 2 * tig + h;
        int row = -1;
        if (t == 0) row = gid;
        else if (t == 1) row = NV + gid;
        else if (gid == 0) row = 8;
        else if (gid < 3) row = NV + NZ + gid - 1;
        if (row >= 0 && col < NV) atomicAdd(&red[row * NV + col], c[t][n][h]);
      }
  __syncthreads();
  for (int e = threadIdx.x; e < NR * NV; e += blockDim.x) atomicAdd(pack + e, red[e]);
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
    // CTAs of 8 warps, 2 resident per SM (98 registers: 12 accumulators + two sets of four prefetched groups), one wave
    const int64_t want_w = ((M + 3) / 4 + 2 * kGramDmmaWarps - 1) / (2 * kGramDmmaWarps);
    const unsigned gd = (unsigned)(want_w < (int64_t)sms * 2 ? (want_w > 0 ? want_w : 1) : (int64_t)sms * 2);
    gram_dmma_kernel<<<gd, kGramDmmaWarps * 32, 0, st>>>(psi, psi_next, u, x, M, pack, seg);
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
