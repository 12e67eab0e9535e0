// percase.cuh -- warp-per-scenario device functions: RLS update, condensed-QP build, exact
// box-QP active-set solve, plant step, small SPD solves.
//
// Mapping: a GROUP of G lanes (G = 32, 16 or 8, template parameter) owns one scenario, so a warp
// carries 32/G scenarios; all per-scenario matrices live in that group's slice of shared memory;
// work inside a phase is distributed over the group's lanes with a lane-strided loop and phases
// are separated by __syncwarp() (all 32 lanes always execute the same phase sequence: groups
// whose scenario index is past the end recompute the last scenario and skip the global writes).
// fp64 throughout (fp32 RLS diverges).
//
// The same source compiles with a host compiler when KMPC_HOSTEMU is defined: a lane-strided
// loop becomes a plain loop over all elements and the warp primitives become no-ops.  That build
// exists only for tests/hostemu (kernel-logic checks on machines without a GPU); it is never
// loaded by the product package.
//
// Reference semantics: RLS duffing.py:927-953, Koopman_update.m:258-278, Tank_System.m:234-263;
// QP Tank_System.m:128-159,182-188 == duffing.py:540-581 + 776-778; plant duffing.py:250-261,
// Koopman_update.m:21-25, Tank_System.m:9-10,211.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/kmpc.h"

#ifdef KMPC_HOSTEMU
inline double rsqrt(double v) { return 1.0 / sqrt(v); }
#define KMPC_DEV inline
#define KMPC_HD
#define KMPC_LANE_LOOP(e, n) for (int e = 0; e < (n); ++e)
#define KMPC_SYNCWARP() ((void)0)
#define KMPC_LANE0 true
#define KMPC_UNROLL
#define KMPC_PDL_WAIT() ((void)0)
#define KMPC_COPY_G2S(dst, src, n) KMPC_LANE_LOOP(e_, n) (dst)[e_] = (src)[e_]
#define KMPC_COPY_G2S_WAIT() ((void)0)
#else
#define KMPC_DEV __device__ __forceinline__
#define KMPC_HD __host__ __device__
// G (lanes per scenario) must be in scope as a compile-time constant
#define KMPC_LANE_LOOP(e, n) for (int e = (int)(threadIdx.x & (G - 1)); e < (n); e += G)
#define KMPC_SYNCWARP() __syncwarp()
#define KMPC_LANE0 ((threadIdx.x & (G - 1)) == 0)
#define KMPC_UNROLL _Pragma("unroll")
// wait for the predecessor kernel, THEN allow the successor to start its own prologue: triggering
// only after the wait keeps the chain transitive (a successor's prologue can overlap this kernel,
// never this kernel's predecessor)
#define KMPC_PDL_WAIT()                                          \
  do {                                                           \
    asm volatile("griddepcontrol.wait;" ::: "memory");           \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  } while (0)
// global -> shared copy of n doubles by the lanes of a group without staging in registers (cp.async, SASS
// LDGSTS): every element of every array is in flight at once, the latency is paid once at the wait
#define KMPC_COPY_G2S(dst, src, n)                                                                     \
  KMPC_LANE_LOOP(e_, n)                                                                                \
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared((dst) + e_)), \
               "l"((src) + e_)                                                                         \
               : "memory")
#define KMPC_COPY_G2S_WAIT() asm volatile("cp.async.wait_all;" ::: "memory")
#endif

namespace kmpc {

// a Cholesky pivot below kPivotFloor * (its original diagonal entry) is replaced by that floor and
// flagged KMPC_STATUS_PIVOT: the Tank Delta-u Hessian reaches cond ~ 2e16 in the RLS transient
constexpr double kPivotFloor = 1e-13;

// ---------------------------------------------------------------- warp reductions -----------
// Reductions over the G lanes of a group (xor offsets < G stay inside the aligned group);
// host build: identity.  (value, index) argmin breaks ties towards the lowest index.
template <int G>
KMPC_DEV void group_argmin(double& val, int& idx) {
#ifndef KMPC_HOSTEMU
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, val, off);
    int oi = __shfl_xor_sync(0xffffffffu, idx, off);
    if (ov < val || (ov == val && oi < idx)) {
      val = ov;
      idx = oi;
    }
  }
#endif
}
template <int G>
KMPC_DEV double group_max(double v) {
#ifndef KMPC_HOSTEMU
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
#endif
  return v;
}
template <int G>
KMPC_DEV int group_or(int v) {
#ifndef KMPC_HOSTEMU
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, off);
#endif
  return v;
}
// true when any lane of the WARP has the predicate set (keeps the groups of a warp in lock step)
KMPC_DEV bool warp_any(bool pred) {
#ifndef KMPC_HOSTEMU
  return __any_sync(0xffffffffu, pred);
#else
  return pred;
#endif
}

// ---------------------------------------------------------------- plant ---------------------
KMPC_DEV void poly2_rhs(double x1, double x2, double u, const double* p, double& d1, double& d2) {
  d1 = p[0] * x2;
  d2 = p[1] * x2 + p[2] * x1 + p[3] * (x1 * x1 * x1) + p[4] * (x1 * x1 * x2) + u;
}

// one plant step for one scenario (scalar code; called by one lane or by every lane redundantly)
KMPC_DEV void plant_step_dev(int kind, int rk4_variant, double h, const double* p, double x1,
                             double x2, double u, double& o1, double& o2) {
  if (kind == KMPC_PLANT_POLY2) {
    double k1a, k1b, k2a, k2b, k3a, k3b, k4a, k4b;
    poly2_rhs(x1, x2, u, p, k1a, k1b);
    poly2_rhs(x1 + 0.5 * h * k1a, x2 + 0.5 * h * k1b, u, p, k2a, k2b);
    poly2_rhs(x1 + 0.5 * h * k2a, x2 + 0.5 * h * k2b, u, p, k3a, k3b);
    if (rk4_variant == KMPC_RK4_PYTHON)
      poly2_rhs(x1 + h * k3a, x2 + h * k3b, u, p, k4a, k4b);
    else  // Koopman_update.m:24: k4 = f(x + k1*dt)
      poly2_rhs(x1 + h * k1a, x2 + h * k1b, u, p, k4a, k4b);
    o1 = x1 + (h / 6.0) * (k1a + 2.0 * k2a + 2.0 * k3a + k4a);
    o2 = x2 + (h / 6.0) * (k1b + 2.0 * k2b + 2.0 * k3b + k4b);
  } else {  // Tank_System.m:9-10 + clamp l.211
    double s1 = sqrt(x1), s2 = sqrt(x2);
    o1 = x1 - p[0] * s1 + p[1] * u;
    o2 = x2 + p[2] * s1 - p[3] * s2;
    if (o1 < 0.0) o1 = 0.0;
    if (o2 < 0.0) o2 = 0.0;
  }
}

// ---------------------------------------------------------------- RBF lift -------------------
KMPC_DEV double rbf_thinplate(const double* x, const double* c, int n, int variant) {
  double r2 = 0.0;
  for (int k = 0; k < n; ++k) {
    double d = x[k] - c[k];
    r2 += d * d;
  }
  if (variant == KMPC_RBF_PYTHON) return r2 * log(sqrt(r2) + 1e-4);  // duffing_RBF.py:22
  if (r2 == 0.0) return 0.0;                                          // rbf.m:26-28 (NaN -> 0)
  return r2 * log(sqrt(r2));
}

// ---------------------------------------------------------------- RLS ------------------------
// Shared-memory workspace of one scenario (doubles).
struct RlsWs {
  double *KA, *P, *barX, *barQ;  // state: nz*nv, nv*nv, n*nz, nz*nz
  double *v, *y, *xc;            // sample: nv (= [z;u]), nz, n
  double *w, *rrow;              // P v, v'P (nv each; reused for bar_Q with nz)
  double *oA, *oB, *oC;          // outputs staged here: nz*nz, nz, n*nz
};
KMPC_HD inline int rls_ws_doubles(int nz, int n) {
  int nv = nz + 1;
  int t = nz * nv + nv * nv + n * nz + nz * nz + nv + nz + n + 2 * nv + nz * nz + nz + n * nz;
  return (t + 1) & ~1;
}
KMPC_DEV RlsWs rls_ws_carve(double* base, int nz, int n) {
  int nv = nz + 1;
  RlsWs w;
  w.KA = base;
  w.P = w.KA + nz * nv;
  w.barX = w.P + nv * nv;
  w.barQ = w.barX + n * nz;
  w.v = w.barQ + nz * nz;
  w.y = w.v + nv;
  w.xc = w.y + nz;
  w.w = w.xc + n;
  w.rrow = w.w + nv;
  w.oA = w.rrow + nv;
  w.oB = w.oA + nz * nz;
  w.oC = w.oB + nz;
  return w;
}

// State and sample already in ws.  Leaves A (nz*nz), B (nz), C (n*nz) in ws.oA / oB / oC.
// Formula order follows the reference (no symmetrisation of P).
template <int G>
KMPC_DEV void rls_update_warp(const RlsWs& ws, int nz, int n, double lam, int flags) {
  const int nv = nz + 1;
  double* oA = ws.oA;
  double* oB = ws.oB;
  double* oC = ws.oC;
  // w = P v (lanes 0..nv-1), rrow = v'P (next nv)
  KMPC_LANE_LOOP(o, 2 * nv) {
    double s = 0.0;
    if (o < nv) {
      KMPC_UNROLL for (int j = 0; j < nv; ++j) s += ws.P[o * nv + j] * ws.v[j];
      ws.w[o] = s;
    } else {
      int j = o - nv;
      KMPC_UNROLL for (int i = 0; i < nv; ++i) s += ws.v[i] * ws.P[i * nv + j];
      ws.rrow[j] = s;
    }
  }
  KMPC_SYNCWARP();
  double vPv = 0.0;  // (v'P) v, duffing.py:934
  KMPC_UNROLL for (int j = 0; j < nv; ++j) vPv += ws.rrow[j] * ws.v[j];
  const double denom = lam + vPv;
  if (lam == 1.0) {  // x / 1 == x exactly: the same bits with one division per entry instead of three
    KMPC_LANE_LOOP(e, nv * nv) {
      int i = e / nv, j = e - i * nv;
      ws.P[e] = ws.P[e] - (ws.w[i] * ws.rrow[j]) / denom;
    }
  } else {           // Koopman_update.m:270 forgetting factor
    KMPC_LANE_LOOP(e, nv * nv) {
      int i = e / nv, j = e - i * nv;
      ws.P[e] = ws.P[e] / lam - (ws.w[i] * ws.rrow[j]) / lam / denom;
    }
  }
  KMPC_LANE_LOOP(e, nz * nv) {
    int i = e / nv, j = e - i * nv;
    ws.KA[e] += ws.y[i] * ws.v[j];
  }
  KMPC_SYNCWARP();
  KMPC_LANE_LOOP(e, nz * nv) {  // [A B] = K_A P
    int i = e / nv, j = e - i * nv;
    double s = 0.0;
    KMPC_UNROLL for (int k = 0; k < nv; ++k) s += ws.KA[i * nv + k] * ws.P[k * nv + j];
    if (j < nz)
      oA[i * nz + j] = s;
    else
      oB[i] = s;
  }
  if (flags & KMPC_RLS_UPDATE_C) {
    KMPC_SYNCWARP();
    KMPC_LANE_LOOP(o, 2 * nz) {  // w = bar_Q z, rrow = z' bar_Q  (z = v[0..nz))
      double s = 0.0;
      if (o < nz) {
        KMPC_UNROLL for (int j = 0; j < nz; ++j) s += ws.barQ[o * nz + j] * ws.v[j];
        ws.w[o] = s;
      } else {
        int j = o - nz;
        KMPC_UNROLL for (int i = 0; i < nz; ++i) s += ws.v[i] * ws.barQ[i * nz + j];
        ws.rrow[j] = s;
      }
    }
    KMPC_SYNCWARP();
    double zQz = 0.0;
    KMPC_UNROLL for (int j = 0; j < nz; ++j) zQz += ws.rrow[j] * ws.v[j];
    const double dq = 1.0 + zQz;
    KMPC_LANE_LOOP(e, nz * nz) {
      int i = e / nz, j = e - i * nz;
      ws.barQ[e] = ws.barQ[e] - (ws.w[i] * ws.rrow[j]) / dq;
    }
    if (!(flags & KMPC_RLS_SKIP_BARX)) {
      KMPC_LANE_LOOP(e, n * nz) {
        int i = e / nz, j = e - i * nz;
        ws.barX[e] += ws.xc[i] * ws.v[j];
      }
    }
    KMPC_SYNCWARP();
    KMPC_LANE_LOOP(e, n * nz) {  // C = bar_X bar_Q
      int i = e / nz, j = e - i * nz;
      double s = 0.0;
      KMPC_UNROLL for (int k = 0; k < nz; ++k) s += ws.barX[i * nz + k] * ws.barQ[k * nz + j];
      oC[e] = s;
    }
  }
  KMPC_SYNCWARP();
}

// ---------------------------------------------------------------- QP -------------------------
// Packed lower-triangular index.
KMPC_HD inline int tri(int i, int j) { return i * (i + 1) / 2 + j; }

struct QpWs {
  double *A, *B, *Cy;   // model: nzq*nzq, nzq, ny*nzq (Cy unused when identity)
  double *z0;           // nzq
  double *VB, *VZ;      // N*nzq each: VB[t] = A^t B, VZ[t] = A^(t+1) z0
  double *g, *e;        // N*ny each (alias VB / VZ when Cy = I)
  double *H, *L;        // packed lower N(N+1)/2 each: H, Cholesky factor of the free block of 2H
  double *invd;         // N
  double *f, *x, *p, *grad, *lb, *ub;  // N each
  int* W;               // N working-set flags: -1 at lower, +1 at upper, 0 free
};
// The Krylov / output arrays (VB, VZ, g, e) are dead once H and f are built, and the factor L is first
// written by the factorisation after that: they share one region (a horizon-50 workspace drops from 32 to
// 24 KB, i.e. 9 instead of 7 resident scenarios per SM; the Tank workspace from 9.6 to 8 KB).
// Layout of the factor L on the device: row i holds its i + 1 entries padded to an even count, so every row
// starts 16-byte aligned and the dot products of the left-looking factorisation read pairs (LDS.128):
// rows 2m and 2m + 1 are 2m + 2 doubles long.  (The host emulator keeps the plain packed triangle; both fit.)
KMPC_HD inline int ltri(int i, int j) {
  const int m = i >> 1;
  return 2 * m * (m + 1) + ((i & 1) ? 2 * m + 2 : 0) + j;
}
// single-slot shapes (N <= 32: Tank) keep the plain packed triangle: their rows are short, the scalar dot
// products win and the cheaper index arithmetic is worth 3 %
template <int NMAX>
KMPC_HD inline int lidx(int i, int j) {
  return NMAX > 32 ? ltri(i, j) : tri(i, j);
}
KMPC_HD inline int qp_ws_build_doubles(int nzq, int ny, int N, bool identity) {
  const int build = 2 * N * nzq + (identity ? 0 : 2 * N * ny), fac = ltri(N, 0);
  return build > fac ? build : fac;
}
KMPC_HD inline int qp_ws_doubles(int nzq, int ny, int N, bool identity) {
  int t = nzq * nzq + nzq + (identity ? 0 : ny * nzq) + nzq + 1 /* L starts 16-byte aligned */ +
          qp_ws_build_doubles(nzq, ny, N, identity) + N * (N + 1) / 2 + 7 * N + (N + 1) / 2;
  return (t + 1) & ~1;  // keep every warp slice 16-byte aligned
}
KMPC_DEV QpWs qp_ws_carve(double* base, int nzq, int ny, int N, bool identity) {
  QpWs w;
  double* p = base;
  w.A = p; p += nzq * nzq;
  w.B = p; p += nzq;
  w.Cy = p; p += identity ? 0 : ny * nzq;
  w.z0 = p; p += nzq;
  p += (p - base) & 1;           // the factor's rows are read in pairs: 16-byte alignment (slices are aligned)
  w.L = p;                       // aliases the build arrays below
  w.VB = p;
  w.VZ = p + N * nzq;
  if (identity) {
    w.g = w.VB;
    w.e = w.VZ;
  } else {
    w.g = p + 2 * N * nzq;
    w.e = p + 2 * N * nzq + N * ny;
  }
  p += qp_ws_build_doubles(nzq, ny, N, identity);
  w.H = p; p += N * (N + 1) / 2;
  w.invd = p; p += N;
  w.f = p; p += N;
  w.x = p; p += N;
  w.p = p; p += N;
  w.grad = p; p += N;
  w.lb = p; p += N;
  w.ub = p; p += N;
  w.W = reinterpret_cast<int*>(p);
  return w;
}

// Build H (packed) and f from the model in ws (A, B, Cy, z0) and the reference r.
//   r_stride = 0: r is (ny) constant over the horizon; r_stride = ny: r is (N, ny).
//   PN: optional terminal weight (ny*ny, row-major) replacing the last q*I block (nullable).
// NZQ > 0: compile-time model dimension -- the lane's row of A stays in registers over the N sequential
// matvec steps (one load per FMA instead of two).
template <int G, int NZQ = 0>
KMPC_DEV void qp_build_warp(const QpWs& ws, int nzq, int ny, int N, bool identity, double q,
                            double rw, const double* r, int r_stride, const double* PN) {
  // ---- Krylov chains: VZ[t] = A VZ[t-1] (VZ[-1] = z0), VB[t+1] = A VB[t] (VB[0] = B);
  //      outputs 0..nzq-1 advance the z chain, nzq..2nzq-1 the B chain, in the same phase
  KMPC_LANE_LOOP(i, nzq) ws.VB[i] = ws.B[i];
  KMPC_SYNCWARP();
#ifndef KMPC_HOSTEMU
  if (NZQ > 0 && 2 * NZQ <= G) {   // one output per lane: lane o < NZQ advances the z chain, NZQ <= o < 2 NZQ the B chain
    const int o = (int)(threadIdx.x & (G - 1));
    const bool half = o >= NZQ, on = o < 2 * NZQ;
    const int i = on ? (half ? o - NZQ : o) : 0;
    double arow[NZQ > 0 ? NZQ : 1];
#pragma unroll
    for (int j = 0; j < NZQ; ++j) arow[j] = ws.A[i * NZQ + j];
    for (int t = 0; t < N; ++t) {
      if (on && (!half || t + 1 < N)) {
        const double* src = half ? ws.VB + t * NZQ : ((t == 0) ? ws.z0 : ws.VZ + (t - 1) * NZQ);
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NZQ; ++j) s += arow[j] * src[j];
        (half ? ws.VB + (t + 1) * NZQ : ws.VZ + t * NZQ)[i] = s;
      }
      __syncwarp();
    }
  } else
#endif
  for (int t = 0; t < N; ++t) {
    KMPC_LANE_LOOP(o, 2 * nzq) {
      const int half = o >= nzq, i = half ? o - nzq : o;
      if (half == 0) {
        const double* src = (t == 0) ? ws.z0 : ws.VZ + (t - 1) * nzq;
        double s = 0.0;
        KMPC_UNROLL for (int j = 0; j < nzq; ++j) s += ws.A[i * nzq + j] * src[j];
        ws.VZ[t * nzq + i] = s;
      } else if (t + 1 < N) {
        const double* src = ws.VB + t * nzq;
        double s = 0.0;
        KMPC_UNROLL for (int j = 0; j < nzq; ++j) s += ws.A[i * nzq + j] * src[j];
        ws.VB[(t + 1) * nzq + i] = s;
      }
    }
    KMPC_SYNCWARP();
  }
  // ---- outputs: g[t] = Cy VB[t], e[t] = Cy VZ[t] - r[t]
  if (identity) {
    KMPC_LANE_LOOP(o, N * ny) {
      int t = o / ny, c = o - t * ny;
      ws.e[o] = ws.VZ[o] - r[t * r_stride + c];
    }
  } else {
    KMPC_LANE_LOOP(o, 2 * N * ny) {
      int which = o / (N * ny), oo = o - which * N * ny;
      int t = oo / ny, c = oo - t * ny;
      const double* src = (which == 0 ? ws.VB : ws.VZ) + t * nzq;
      double s = 0.0;
      KMPC_UNROLL for (int j = 0; j < nzq; ++j) s += ws.Cy[c * nzq + j] * src[j];
      if (which == 0)
        ws.g[oo] = s;
      else
        ws.e[oo] = s - r[t * r_stride + c];
    }
  }
  KMPC_SYNCWARP();
  // ---- H along diagonals: H[a][a-d] = q * sum_{t=0}^{N-1-a} <g[t+d], g[t]>  (+ rw on d = 0)
  KMPC_LANE_LOOP(d, N) {
    double run = 0.0;
    for (int t = 0; t + d < N; ++t) {
      double s = 0.0;
      KMPC_UNROLL for (int c = 0; c < ny; ++c) s += ws.g[(t + d) * ny + c] * ws.g[t * ny + c];
      run += q * s;
      int a = N - 1 - t;
      ws.H[tri(a, a - d)] = run + (d == 0 ? rw : 0.0);
    }
  }
  // ---- f[a] = 2 q sum_{k=a}^{N-1} <g[k-a], e[k]>
  KMPC_LANE_LOOP(a, N) {
    double s = 0.0;
    for (int k = a; k < N; ++k)
      KMPC_UNROLL for (int c = 0; c < ny; ++c) s += ws.g[(k - a) * ny + c] * ws.e[k * ny + c];
    ws.f[a] = 2.0 * q * s;
  }
  KMPC_SYNCWARP();
  if (PN != nullptr) {
    // terminal block: H[a][b] += g[N-1-a]' (PN - qI) g[N-1-b],  f[a] += 2 g[N-1-a]' (PN - qI) e[N-1]
    KMPC_LANE_LOOP(a, N) {
      double s = 0.0;
      for (int c = 0; c < ny; ++c)
        for (int k = 0; k < ny; ++k)
          s += ws.g[(N - 1 - a) * ny + c] * (PN[c * ny + k] - (c == k ? q : 0.0)) *
               ws.e[(N - 1) * ny + k];
      ws.f[a] += 2.0 * s;
    }
    KMPC_LANE_LOOP(pr, N * (N + 1) / 2) {
      int a = (int)((sqrt(8.0 * pr + 1.0) - 1.0) * 0.5);
      while (tri(a + 1, 0) <= pr) ++a;
      while (tri(a, 0) > pr) --a;
      const int b = pr - tri(a, 0);
      double s = 0.0;
      for (int c = 0; c < ny; ++c)
        for (int k = 0; k < ny; ++k)
          s += ws.g[(N - 1 - a) * ny + c] * (PN[c * ny + k] - (c == k ? q : 0.0)) *
               ws.g[(N - 1 - b) * ny + k];
      ws.H[pr] += s;
    }
    KMPC_SYNCWARP();
  }
}

#ifndef KMPC_HOSTEMU
// 1 / sqrt(d) for a positive normal d (the pivots are floored above zero; NaN propagates): hardware
// seed (MUFU.RSQ64H, ~2^-26 relative) and one third-order correction y (1 + e/2 + 3 e^2/8), e = 1 - d y^2,
// which leaves ~2^-78 before the final rounding -- the accuracy of rsqrt() without its special-case paths.
__device__ __forceinline__ double rsqrt_pos(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double t = d * y;
  const double e = fma(-t, y, 1.0);
  const double c = fma(0.375, e, 0.5);
  return fma(y, e * c, y);
}

// ---- GPU versions of the factorisation, the triangular solves and the gradient -----------------
// Lane l of the group OWNS rows l, l + G, ... of the packed lower triangle (slots; N <= 64 gives
// at most 64 / G of them).  A row's entries are only ever written by its owner, so
//   * the left-looking column step needs ONE warp synchronisation per column (row j, read by
//     everybody in later columns), the pivot travels by shuffle, the dot products run on four
//     independent accumulators over row pointers hoisted out of the loop (the triangular numbers
//     T(i) mod 16 are a permutation of 0..15 over any 16 consecutive i, so lanes reading
//     L[T(i) + k] hit distinct banks);
//   * the triangular solves keep the right-hand side in registers and broadcast y_j / x_j with a
//     shuffle: no synchronisation at all.
// The host emulator (tests/hostemu) runs the lane-loop versions below; results agree to rounding.
constexpr int kQpSlotsMax = KMPC_MAX_HORIZON / 16;

// Free-set compaction.  The working set of a Tank / horizon-50 QP in a transient holds most of the
// variables (measured on the Tank loop: 7.9 free of 20 on average over all factorisations), so the
// factorisation and the triangular solves run on the COMPACTED free block: compact row c <-> original
// variable orig(c), increasing.  Every lane of the group knows the free mask (ballot), lane c owns
// compact rows c, c + G, ... and carries their original indices; the column owner broadcasts its
// original index by shuffle.  ws.L / ws.invd hold the compact factor (nf(nf+1)/2 entries).
// Loop trip counts use the maximum nf over the groups of the warp (the shuffles are warp-wide).
// NMAX: compile-time upper bound of the horizon (sizes the per-lane slots: a Tank QP, N = 20 on 32 lanes,
// needs one slot, not the 64 / G of the run-time shape -- half the code, which matters because the single-warp
// blocks of this kernel sit at unrelated program counters and live off the instruction cache).
template <int G, int NMAX = KMPC_MAX_HORIZON>
struct QpFreeMap {
  static constexpr int SLOTS = (NMAX + G - 1) / G;
  int nf, nfw;       // free variables of this group / maximum over the groups of the warp
  int oi[SLOTS];     // original index of compact row lane + sl * G (-1 beyond nf)
  __device__ __forceinline__ void build(const QpWs& ws, int N) {
    const int lane = threadIdx.x & (G - 1);
    // every free variable knows its rank among the free ones (ballot + popc below its lane) and scatters its
    // index to that slot of a small map -- ws.invd is dead until the factorisation that follows writes it --,
    // lane c then reads the c-th free index back.  (Searching the n-th set bit with __fns was 6 % of the Tank
    // kernel's instructions.)
    int* cmap = reinterpret_cast<int*>(ws.invd);
    int cum = 0;
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {
      const int i = lane + sl * G;
      const bool fr = (i < N) && (ws.W[i] == 0);
      unsigned b = __ballot_sync(0xffffffffu, fr);
      if (G < 32) b = (b >> (threadIdx.x & 31 & ~(G - 1))) & ((1u << G) - 1u);
      if (fr) cmap[cum + __popc(b & ((1u << lane) - 1u))] = i;
      cum += __popc(b);
    }
    nf = cum;
    nfw = nf;
    if (G < 32) {
#pragma unroll
      for (int o = G; o < 32; o <<= 1) nfw = max(nfw, __shfl_xor_sync(0xffffffffu, nfw, o));
    }
    __syncwarp();
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {
      const int c = lane + sl * G;
      oi[sl] = (c < nf) ? cmap[c] : -1;
    }
    __syncwarp();   // the map has been read: the factorisation may write ws.invd
  }
};

// Horizon-50 shape (one scenario per warp, two row slots per lane): the k < j0 part of the dot products of a
// whole panel of 8 columns j0 .. j0 + 7 is ONE small GEMM, S = L[j0.., 0..j0) L[j0..j0+8, 0..j0)', done on the fp64
// tensor path before the panel's columns are eliminated: a DMMA (m8n8k4) does the work of 256 scalar FMAs with
// two loads, its fragments come straight from the row-major factor (lane (gid, tig) reads L[row0 + gid][k0 + tig]),
// and S is parked in the panel's own -- not yet written -- entries of L.  The column steps then only add the
// k in [j0, j) terms.  45 % of this kernel's samples were those dot products.
__device__ __forceinline__ void qp_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
template <int NMAX>
__device__ __forceinline__ void qp_chol_panel_dmma(const QpWs& ws, int nf, int j0) {
  constexpr int MAXB = (NMAX + 7) / 8;
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const int nb = (nf + 7) >> 3, jb = j0 >> 3;
  double c[MAXB][2];
#pragma unroll
  for (int t = 0; t < MAXB; ++t) c[t][0] = c[t][1] = 0.0;
  const int rb = j0 + gid;
  const double* prow = ws.L + lidx<NMAX>(rb < nf ? rb : 0, tig);
  for (int k0 = 0; k0 < j0; k0 += 4) {
    const double b = (rb < nf) ? prow[k0] : 0.0;
    qp_dmma(c[0][0], c[0][1], b, b);          // the panel's own row block: A = B
#pragma unroll
    for (int t = 1; t < MAXB; ++t) {
      if (jb + t < nb) {                        // warp-uniform
        const int ra = 8 * (jb + t) + gid;
        const double a = (ra < nf) ? ws.L[lidx<NMAX>(ra, k0 + tig)] : 0.0;
        qp_dmma(c[t][0], c[t][1], a, b);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < MAXB; ++t) {
    if (jb + t < nb) {
      const int row = 8 * (jb + t) + gid;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = j0 + 2 * tig + h;
        if (row < nf && col <= row) ws.L[lidx<NMAX>(row, col)] = c[t][h];
      }
    }
  }
  __syncwarp();
}

template <int G, int NMAX>
__device__ __forceinline__ int qp_chol_masked_rows(const QpWs& ws, int N, const QpFreeMap<G, NMAX>& fm) {
  constexpr int SLOTS = QpFreeMap<G, NMAX>::SLOTS;
  const int lane = threadIdx.x & (G - 1);
  int status = 0;
  const int nf = fm.nf;
  int rbase[SLOTS];   // offset of this lane's rows in the factor (the index arithmetic was 10 % of the instructions)
#pragma unroll
  for (int sl = 0; sl < SLOTS; ++sl) rbase[sl] = lidx<NMAX>(lane + sl * G, 0);
#pragma unroll
  for (int so = 0; so < SLOTS; ++so) {
    for (int j = so * G; j < fm.nfw && j < (so + 1) * G; ++j) {
      const int oj = __shfl_sync(0xffffffffu, fm.oi[so], j & (G - 1), G);   // -1 when j >= nf (another group's column)
      const bool live = j < nf;
      constexpr bool PANELS = (G == 32) && (NMAX > 32) && (NMAX < KMPC_MAX_HORIZON);
      const int kbeg = PANELS ? (j & ~7) : 0;   // first k the column step still has to add
      if (PANELS && kbeg == j && j > 0 && live) qp_chol_panel_dmma<NMAX>(ws, nf, j);
      // dot products of row j with the rows below it, four accumulators per row (k mod 4, the tail into the
      // first: the summation order of every version of this kernel); row j's pairs are loaded once for all slots
      const int jbase = lidx<NMAX>(j, 0);
      const double* rowj = ws.L + jbase + kbeg;
      const double2* rj2 = reinterpret_cast<const double2*>(rowj);
      const int jlen = j - kbeg;
      double sv[SLOTS], dmine = 1.0, acc[SLOTS][4];
      bool act[SLOTS];
      const double2* ri2[SLOTS];
#pragma unroll
      for (int sl = 0; sl < SLOTS; ++sl) {
        const int i = lane + sl * G;
        act[sl] = live && i >= j && i < nf;
        ri2[sl] = reinterpret_cast<const double2*>(ws.L + (act[sl] ? rbase[sl] : jbase) + kbeg);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[sl][q] = 0.0;
        if (PANELS && kbeg > 0 && act[sl]) acc[sl][0] = ws.L[rbase[sl] + j];   // S of the panel pre-pass
      }
      if (live && SLOTS == 1) {
        // short rows (Tank: 20 columns, 10 on average): scalar loads on four accumulators measured faster
        // than the paired form (56.7 against 52.9 M scenario-steps/s on the Tank loop)
        if (act[0]) {
          const double* rowi = reinterpret_cast<const double*>(ri2[0]);
          int k = 0;
          for (; k + 3 < jlen; k += 4) {
            acc[0][0] = fma(rowi[k], rowj[k], acc[0][0]);
            acc[0][1] = fma(rowi[k + 1], rowj[k + 1], acc[0][1]);
            acc[0][2] = fma(rowi[k + 2], rowj[k + 2], acc[0][2]);
            acc[0][3] = fma(rowi[k + 3], rowj[k + 3], acc[0][3]);
          }
          for (; k < jlen; ++k) acc[0][0] = fma(rowi[k], rowj[k], acc[0][0]);
        }
      } else if (live) {
        const int blocks = jlen >> 2;
        for (int bq = 0; bq < blocks; ++bq) {
          const double2 b0 = rj2[2 * bq], b1 = rj2[2 * bq + 1];
#pragma unroll
          for (int sl = 0; sl < SLOTS; ++sl) {
            if (act[sl]) {
              const double2 a0 = ri2[sl][2 * bq], a1 = ri2[sl][2 * bq + 1];
              acc[sl][0] = fma(a0.x, b0.x, acc[sl][0]);
              acc[sl][1] = fma(a0.y, b0.y, acc[sl][1]);
              acc[sl][2] = fma(a1.x, b1.x, acc[sl][2]);
              acc[sl][3] = fma(a1.y, b1.y, acc[sl][3]);
            }
          }
        }
        for (int k = blocks << 2; k < jlen; ++k) {
          const double bk = rowj[k];
#pragma unroll
          for (int sl = 0; sl < SLOTS; ++sl)
            if (act[sl]) acc[sl][0] = fma(reinterpret_cast<const double*>(ri2[sl])[k], bk, acc[sl][0]);
        }
      }
#pragma unroll
      for (int sl = 0; sl < SLOTS; ++sl) {
        const int i = lane + sl * G;
        double v = 0.0;
        if (act[sl]) {
          v = 2.0 * ws.H[tri(fm.oi[sl], oj)] - ((acc[sl][0] + acc[sl][1]) + (acc[sl][2] + acc[sl][3]));
          if (i == j) dmine = v;
        }
        sv[sl] = v;
      }
      double d = __shfl_sync(0xffffffffu, dmine, j & (G - 1), G);
      const double floor_j = live ? kPivotFloor * (2.0 * ws.H[tri(oj, oj)]) : 0.0;
      if (live && !(d > floor_j)) {  // numerically semi-definite: regularise and flag (oracle/mpc.py PIVOT_FLOOR)
        status |= KMPC_STATUS_PIVOT;
        d = floor_j;
      }
      const double inv = rsqrt_pos(d);
      if (live) {
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) {
          const int i = lane + sl * G;
          if (i == j) {
            ws.L[jbase + j] = d * inv;
            ws.invd[j] = inv;
          } else if (i > j && i < nf) {
            ws.L[rbase[sl] + j] = sv[sl] * inv;
          }
        }
      }
      __syncwarp();
    }
  }
  return status;
}

template <int G, int NMAX>
__device__ __forceinline__ void qp_chol_solve_rows(const QpWs& ws, int N, const QpFreeMap<G, NMAX>& fm) {
  constexpr int SLOTS = QpFreeMap<G, NMAX>::SLOTS;
  const int lane = threadIdx.x & (G - 1);
  const int nf = fm.nf;
  double pr[SLOTS];
  int rbase[SLOTS];
#pragma unroll
  for (int sl = 0; sl < SLOTS; ++sl) {
    pr[sl] = (lane + sl * G < nf) ? ws.p[fm.oi[sl]] : 0.0;
    rbase[sl] = lidx<NMAX>(lane + sl * G, 0);
  }
  // forward: y_j = p_j / L_jj, then p_i -= L_ij y_j for the rows below
#pragma unroll
  for (int so = 0; so < SLOTS; ++so) {
    for (int j = so * G; j < fm.nfw && j < (so + 1) * G; ++j) {
      const bool live = j < nf;
      const double yj = __shfl_sync(0xffffffffu, pr[so], j & (G - 1), G) * (live ? ws.invd[j] : 0.0);
      if (live) {
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) {
          const int i = lane + sl * G;
          if (i > j && i < nf) pr[sl] = fma(-ws.L[rbase[sl] + j], yj, pr[sl]);
        }
        if (lane == (j & (G - 1))) pr[so] = yj;
      }
    }
  }
  // backward: x_j = y_j / L_jj, then y_i -= L_ji x_j for the rows above (row j of L: coalesced)
#pragma unroll
  for (int so = SLOTS - 1; so >= 0; --so) {
    const int jhi = (fm.nfw < (so + 1) * G ? fm.nfw : (so + 1) * G) - 1;
    for (int j = jhi; j >= so * G; --j) {
      const bool live = j < nf;
      const double xj = __shfl_sync(0xffffffffu, pr[so], j & (G - 1), G) * (live ? ws.invd[j] : 0.0);
      if (live) {
        const double* rowj = ws.L + lidx<NMAX>(j, 0);
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) {
          const int i = lane + sl * G;
          if (i < j) pr[sl] = fma(-rowj[i], xj, pr[sl]);
        }
        if (lane == (j & (G - 1))) pr[so] = xj;
      }
    }
  }
#pragma unroll
  for (int sl = 0; sl < SLOTS; ++sl)
    if (lane + sl * G < nf) ws.p[fm.oi[sl]] = pr[sl];
  __syncwarp();
}

// Compile-time horizon bound (Tank: 20, horizon 50): branch-free walk over the packed triangle --
// H[i][j] sits at T(i) + j for j <= i and at T(j) + i above the diagonal, T(j) a constant after unrolling.
template <int G, int NMAX>
__device__ __forceinline__ void qp_gradient_rows_ct(const QpWs& ws, int N) {
  const int lane = threadIdx.x & (G - 1);
#pragma unroll
  for (int sl = 0; sl < (NMAX + G - 1) / G; ++sl) {
    const int i = lane + sl * G;
    if (i < N) {
      const int ti = tri(i, 0);
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int j = 0; j < NMAX; ++j) {
        if (j < N) {
          const double h = ws.H[(i >= j) ? ti + j : (j * (j + 1)) / 2 + i];
          if (j & 1) s1 = fma(h, ws.x[j], s1);
          else s0 = fma(h, ws.x[j], s0);
        }
      }
      ws.grad[i] = 2.0 * (s0 + s1) + ws.f[i];
    }
  }
  __syncwarp();
}

template <int G>
__device__ __forceinline__ void qp_gradient_rows(const QpWs& ws, int N) {
  const int lane = threadIdx.x & (G - 1);
  for (int i = lane; i < N; i += G) {
    const double* rowi = ws.H + tri(i, 0);
    double s0 = 0.0, s1 = 0.0;
    int j = 0;
    for (; j + 1 <= i; j += 2) {
      s0 = fma(rowi[j], ws.x[j], s0);
      s1 = fma(rowi[j + 1], ws.x[j + 1], s1);
    }
    if (j <= i) s0 = fma(rowi[j], ws.x[j], s0);
    const double* col = ws.H + tri(i + 1, 0) + i;   // H[j][i], j = i + 1 ..: stride j + 1
    for (j = i + 1; j < N; ++j) {
      s1 = fma(col[0], ws.x[j], s1);
      col += j + 1;
    }
    ws.grad[i] = 2.0 * (s0 + s1) + ws.f[i];
  }
  __syncwarp();
}
#endif  // !KMPC_HOSTEMU

#ifdef KMPC_HOSTEMU
// Host-emulator forms (lane loops).  The device goes through qp_factor_solve below: ONE free-set map for the
// factorisation and both solves (the map is staged in ws.invd, which the factorisation then overwrites, so the
// two halves cannot be called separately there).
// Cholesky of the free block of 2H into ws.L (masked rows/cols become identity rows).
// Returns KMPC_STATUS_PIVOT if a pivot was not positive.
template <int G>
KMPC_DEV int qp_chol_masked(const QpWs& ws, int N) {
  int status = 0;
  for (int j = 0; j < N; ++j) {
    const bool mj = ws.W[j] != 0;  // warp-uniform
    KMPC_LANE_LOOP(ii, N - j) {
      int i = j + ii;
      double s = 0.0;
      if (!mj && ws.W[i] == 0) {
        s = 2.0 * ws.H[tri(i, j)];
        for (int k = 0; k < j; ++k) s -= ws.L[tri(i, k)] * ws.L[tri(j, k)];
      }
      ws.L[tri(i, j)] = s;  // unscaled column
    }
    KMPC_SYNCWARP();
    double d = mj ? 1.0 : ws.L[tri(j, j)];
    const double floor_j = mj ? 0.0 : kPivotFloor * (2.0 * ws.H[tri(j, j)]);
    if (!(d > floor_j)) {  // numerically semi-definite: regularise and flag (oracle/mpc.py PIVOT_FLOOR)
      status |= KMPC_STATUS_PIVOT;
      d = floor_j;
    }
    const double inv = rsqrt(d);  // one MUFU + Newton instead of sqrt followed by a division
    const double piv = d * inv;
    KMPC_SYNCWARP();  // everyone has read L[j][j] before it is overwritten
    KMPC_LANE_LOOP(ii, N - j) {
      int i = j + ii;
      if (ii == 0) {
        ws.L[tri(j, j)] = piv;
        ws.invd[j] = inv;
      } else {
        ws.L[tri(i, j)] *= inv;
      }
    }
    KMPC_SYNCWARP();
  }
  return status;
}

// Solve (L L') p = rhs in place in ws.p (rhs must be zero on masked entries).
template <int G>
KMPC_DEV void qp_chol_solve(const QpWs& ws, int N) {
  for (int j = 0; j < N; ++j) {  // forward, column oriented
    const double yj = ws.p[j] * ws.invd[j];
    KMPC_SYNCWARP();
    KMPC_LANE_LOOP(ii, N - j) {
      int i = j + ii;
      if (ii == 0)
        ws.p[j] = yj;
      else
        ws.p[i] -= ws.L[tri(i, j)] * yj;
    }
    KMPC_SYNCWARP();
  }
  for (int j = N - 1; j >= 0; --j) {  // backward
    const double xj = ws.p[j] * ws.invd[j];
    KMPC_SYNCWARP();
    KMPC_LANE_LOOP(i, j + 1) {
      if (i == j)
        ws.p[j] = xj;
      else
        ws.p[i] -= ws.L[tri(j, i)] * xj;
    }
    KMPC_SYNCWARP();
  }
}
#endif  // KMPC_HOSTEMU

// grad = 2 H x + f
template <int G, int NMAX = KMPC_MAX_HORIZON>
KMPC_DEV void qp_gradient(const QpWs& ws, int N) {
#ifndef KMPC_HOSTEMU
  if (NMAX < KMPC_MAX_HORIZON) qp_gradient_rows_ct<G, NMAX>(ws, N);
  else qp_gradient_rows<G>(ws, N);
#else
  KMPC_LANE_LOOP(i, N) {
    double s = 0.0;
    for (int j = 0; j < N; ++j) s += ws.H[i >= j ? tri(i, j) : tri(j, i)] * ws.x[j];
    ws.grad[i] = 2.0 * s + ws.f[i];
  }
  KMPC_SYNCWARP();
#endif
}

// ws.p <- (2H)_FF^-1 ws.p on the free block F = {W == 0}: factorisation + both triangular solves with ONE
// free-set map (W does not change in between).  Returns the factorisation's status bits.
template <int G, int NMAX>
KMPC_DEV int qp_factor_solve(const QpWs& ws, int N) {
#ifndef KMPC_HOSTEMU
  QpFreeMap<G, NMAX> fm;
  fm.build(ws, N);
  const int st = qp_chol_masked_rows<G, NMAX>(ws, N, fm);
  qp_chol_solve_rows<G, NMAX>(ws, N, fm);
  return st;
#else
  const int st = qp_chol_masked<G>(ws, N);
  qp_chol_solve<G>(ws, N);
  return st;
#endif
}

// Exact solve of  min x'Hx + f'x,  lb <= x <= ub  (oracle/mpc.py solve_box_qp_exact is the same
// algorithm).  Start: clipped unconstrained minimiser, clipped variables in the working set.
// Iterations 0..kPdasIters-1 are primal-dual active-set sweeps (full Newton step on the free
// face, then clip EVERY violated bound and release EVERY bound with a negative multiplier; stop
// when nothing changes: KKT holds exactly); later iterations, reached only if the sweeps cycle,
// are the monotone primal active-set method (ratio test; add the blocking bound or drop the most
// negative multiplier), which always terminates.  Result in ws.x; returns status bits.
// The groups of a warp iterate in lock step: a group that has converged keeps executing the
// phases (its x and W are frozen) until every group of the warp is done.
constexpr int kPdasIters = 8;
// The warp-per-scenario solver (Tank, horizon 50) damps the sweeps instead of giving up on them: from sweep
// kPdasReleaseAll on, a sweep still clips EVERY violated bound but releases only the bound with the most
// negative multiplier.  Undamped sweeps cycle on these Hessians in more than half of the heavy steps and
// the monotone fallback then pays one factorisation per bound (up to 59 measured on the Tank workload);
// damped, they converge within kPdasItersDamped sweeps in all but 0.5 % of the steps (-18 % factorisations).
constexpr int kPdasItersDamped = 40;
constexpr int kPdasReleaseAll = 1;

// `warm` (uniform over the warp): ws.x already holds a start point -- the previous step's optimal move
// sequence shifted by one move (receding horizon).  It is clipped into the box, every variable at
// a bound enters the working set, and the monotone primal method runs from there: when the optimal
// set moves by a few bounds per step this costs a few factorisations, where a cold start needs
// one per sweep and then one per bound of the final set whenever the sweeps cycle (Tank: 10 - 25).
template <int G, int NMAX = KMPC_MAX_HORIZON>
KMPC_DEV int qp_solve_warp(const QpWs& ws, int N, int max_iter, double tol, bool warm = false,
                           bool warm_sweeps = true, bool damped = false) {
  int status = 0;
  int any = 0;
  double fmaxabs = 0.0;
  if (!warm) {
    KMPC_LANE_LOOP(i, N) {
      ws.W[i] = 0;
      ws.p[i] = -ws.f[i];
    }
    KMPC_SYNCWARP();
    status |= qp_factor_solve<G, NMAX>(ws, N);
    KMPC_LANE_LOOP(i, N) {
      double xi = ws.p[i];
      int w = 0;
      if (xi < ws.lb[i]) {
        w = -1;
        xi = ws.lb[i];
      } else if (xi > ws.ub[i]) {
        w = 1;
        xi = ws.ub[i];
      }
      ws.W[i] = w;
      ws.x[i] = xi;
      any |= (w != 0);
      fmaxabs = fmax(fmaxabs, fabs(ws.f[i]));
    }
    any = group_or<G>(any);
  } else {
    KMPC_LANE_LOOP(i, N) {
      const double xi = fmin(fmax(ws.x[i], ws.lb[i]), ws.ub[i]);
      ws.x[i] = xi;
      ws.W[i] = xi <= ws.lb[i] ? -1 : (xi >= ws.ub[i] ? 1 : 0);
      fmaxabs = fmax(fmaxabs, fabs(ws.f[i]));
    }
    any = 1;
  }
  fmaxabs = group_max<G>(fmaxabs);
  KMPC_SYNCWARP();
  const double mtol = tol * fmax(1.0, fmaxabs);
  bool done = !any;
  if (warp_any(!done)) qp_gradient<G, NMAX>(ws, N);  // grad at the start point; refreshed after each step
  for (int it = 0; it < max_iter; ++it) {
    if (!warp_any(!done)) break;
    // primal-dual sweeps first, also from a warm working set (a good guess converges in 1 - 3 sweeps where
    // the primal method needs one factorisation per changed bound); warm_sweeps = false: primal only
    const bool pdas = (!warm || warm_sweeps) && it < (damped ? kPdasItersDamped : kPdasIters);  // uniform over the warp
    KMPC_LANE_LOOP(i, N) ws.p[i] = (ws.W[i] == 0) ? -ws.grad[i] : 0.0;
    KMPC_SYNCWARP();
    const int cst = qp_factor_solve<G, NMAX>(ws, N);
    if (!done) status |= cst;
    double alpha = 1.0;
    int block = 0x7fffffff;
    if (!pdas) {  // ratio test
      KMPC_LANE_LOOP(i, N) {
        if (ws.W[i] == 0) {
          double pi = ws.p[i], xi = ws.x[i], a = 2.0;
          if (pi > 0.0 && xi + pi > ws.ub[i])
            a = (ws.ub[i] - xi) / pi;
          else if (pi < 0.0 && xi + pi < ws.lb[i])
            a = (ws.lb[i] - xi) / pi;
          if (a < alpha) {  // strict: lowest index wins ties within a lane's ascending sweep
            alpha = a;
            block = i;
          }
        }
      }
      group_argmin<G>(alpha, block);
    }
    const bool blocked = block != 0x7fffffff;
    if (!done) {
      KMPC_LANE_LOOP(i, N) {
        double xi = ws.x[i] + alpha * ws.p[i];
        if (blocked && i == block) {
          const int side = ws.p[i] > 0.0 ? 1 : -1;
          xi = side > 0 ? ws.ub[i] : ws.lb[i];
          ws.W[i] = side;
        }
        ws.x[i] = xi;
      }
    }
    KMPC_SYNCWARP();
    qp_gradient<G, NMAX>(ws, N);  // at the new point (PDAS: the unclipped face minimiser)
    if (pdas) {
      int changed = 0, clipped = 0;
      const bool single = damped && it >= kPdasReleaseAll;   // uniform over the warp
      double worst = INFINITY;
      int widx = 0x7fffffff;
      if (!done) {
        KMPC_LANE_LOOP(i, N) {
          const int w = ws.W[i];
          if (w != 0) {  // release every bound whose multiplier has the wrong sign (damped: only the worst)
            const double lam = w < 0 ? ws.grad[i] : -ws.grad[i];
            if (lam < -mtol) {
              if (!single) {
                ws.W[i] = 0;
                changed = 1;
              } else if (lam < worst) {
                worst = lam;
                widx = i;
              }
            }
          } else {       // clip every violated bound into the working set
            const double xi = ws.x[i];
            if (xi < ws.lb[i]) {
              ws.x[i] = ws.lb[i];
              ws.W[i] = -1;
              clipped = 1;
            } else if (xi > ws.ub[i]) {
              ws.x[i] = ws.ub[i];
              ws.W[i] = 1;
              clipped = 1;
            }
          }
        }
      }
      if (single) {
        group_argmin<G>(worst, widx);
        if (!done && widx != 0x7fffffff) {
          if (KMPC_LANE0) ws.W[widx] = 0;
          changed = 1;
        }
      }
      clipped = group_or<G>(clipped);
      changed = group_or<G>(changed) | clipped;
      if (!done && !changed) done = true;
      KMPC_SYNCWARP();
      if (warp_any(clipped != 0)) qp_gradient<G, NMAX>(ws, N);
    } else {
      // after a full step: multipliers of the bound variables
      double worst = INFINITY;
      int widx = 0x7fffffff;
      KMPC_LANE_LOOP(i, N) {
        const int w = ws.W[i];
        if (w != 0) {
          const double lam = w < 0 ? ws.grad[i] : -ws.grad[i];
          if (lam < worst) {
            worst = lam;
            widx = i;
          }
        }
      }
      group_argmin<G>(worst, widx);
      if (!done && !blocked) {
        if (widx == 0x7fffffff || worst >= -mtol) {
          done = true;
        } else if (KMPC_LANE0) {
          ws.W[widx] = 0;
        }
      }
      KMPC_SYNCWARP();
    }
  }
  if (!done) status |= KMPC_STATUS_MAXITER;
  int bad = 0;
  KMPC_LANE_LOOP(i, N) bad |= !isfinite(ws.x[i]);
  if (group_or<G>(bad)) status |= KMPC_STATUS_NONFINITE;
  return status;
}


// ---------------------------------------------------------------- small SPD solve ------------
// X = Bm * inv(G) for SPD G (n x n, row-major, destroyed) and Bm (rows x n): Cholesky of G then
// two triangular solves per row.  Used by the EDMD solve (K = Aq G^-1).  Work arrays in smem or
// global; single warp.
template <int G>
KMPC_DEV int spd_right_solve_warp(double* Gm, int n, double* Bm, int rows) {
  int status = 0;
  for (int j = 0; j < n; ++j) {  // in-place lower Cholesky, left-looking
    KMPC_LANE_LOOP(ii, n - j) {
      int i = j + ii;
      double s = Gm[i * n + j];
      for (int k = 0; k < j; ++k) s -= Gm[i * n + k] * Gm[j * n + k];
      Gm[i * n + j] = s;
    }
    KMPC_SYNCWARP();
    double d = Gm[j * n + j];
    double orig = d;  // diagonal before elimination = d + sum_k L[j][k]^2
    for (int k = 0; k < j; ++k) orig += Gm[j * n + k] * Gm[j * n + k];
    if (!(d > 1e-12 * orig)) {  // (numerically) rank deficient: pinv and inverse differ, flag it
      status |= KMPC_STATUS_PIVOT;
      if (!(d > 0.0)) d = 1e-300;
    }
    const double piv = sqrt(d);
    KMPC_SYNCWARP();
    KMPC_LANE_LOOP(ii, n - j) {
      int i = j + ii;
      Gm[i * n + j] = (ii == 0) ? piv : Gm[i * n + j] / piv;
    }
    KMPC_SYNCWARP();
  }
  // each row b of Bm solves  x (L L') = b  ->  L L' x' = b'
  KMPC_LANE_LOOP(rr, rows) {
    double* b = Bm + rr * n;
    for (int i = 0; i < n; ++i) {
      double s = b[i];
      for (int k = 0; k < i; ++k) s -= Gm[i * n + k] * b[k];
      b[i] = s / Gm[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = b[i];
      for (int k = i + 1; k < n; ++k) s -= Gm[k * n + i] * b[k];
      b[i] = s / Gm[i * n + i];
    }
  }
  KMPC_SYNCWARP();
  return status;
}

}  // namespace kmpc
