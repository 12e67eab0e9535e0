// loopbody.cuh -- per-scenario bodies of the fused closed-loop kernels (one warp per scenario).
// Written with the percase.cuh lane-loop macros so tests/hostemu can run the identical code
// sequentially on the host.  See closed_loop.cu for the kernels and the step schedule.
#pragma once
#include "percase.cuh"

namespace kmpc {

struct LoopDev {  // passed by value to the kernels
  kmpc_loop_config c;
  kmpc_loop_buffers b;
  double* z_next;  // ctx-owned (S, nz): lift(x+)
  double* x_prev;  // ctx-owned (S, n):  x before the plant step (tank C regression)
  unsigned int* wset;  // ctx-owned (S, 2): optimal QP working set of the last step (fused kernel warm start)
  double* qp_x;        // ctx-owned (S, N), nullable: optimal move sequence of the last step (generic kernels' warm
                       // start); NaN in move 0 = no warm start for this scenario
};

KMPC_HD inline int loop_nzq(const kmpc_loop_config& c) { return c.nz + (c.du_aug ? 1 : 0); }
KMPC_HD inline int loop_ny(const kmpc_loop_config& c) {
  return c.out_mode == KMPC_OUT_IDENTITY ? loop_nzq(c) : (c.out_mode == KMPC_OUT_C ? c.n : 1);
}

// Shape of the QP a loop solves; passed explicitly (not read from the config) so that kernel
// instantiations with compile-time shapes get their loops unrolled and indices folded.
struct LoopShape {
  int nz, N, out_mode, du_aug;
};
KMPC_HD inline LoopShape loop_shape(const kmpc_loop_config& c) {
  LoopShape sh = {c.nz, c.N, c.out_mode, c.du_aug ? 1 : 0};
  return sh;
}

// QP + plant for scenario s at closed-loop step `step`; `base` = this group's smem slice;
// `valid` = false for the padding groups of the last warp (compute, but write nothing).
template <int G, int NMAX = KMPC_MAX_HORIZON, int NZQ = 0>
KMPC_DEV void loop_qp_plant_scenario(const LoopDev& d, const LoopShape sh, int64_t s, bool valid,
                                     int64_t step, int64_t log_slot, double* base) {
  const kmpc_loop_config& c = d.c;
  const int nz = sh.nz, n = 2, N = sh.N;
  const int nzq = nz + sh.du_aug;
  const bool identity = sh.out_mode == KMPC_OUT_IDENTITY;
  const int ny = identity ? nzq : (sh.out_mode == KMPC_OUT_C ? n : 1);
  QpWs ws = qp_ws_carve(base, nzq, ny, N, identity);
  const int64_t sm = c.shared_model ? 0 : s;
  const double* A = d.b.A + sm * nz * nz;
  const double* B = d.b.B + sm * nz;
  const double* C = d.b.C + sm * n * nz;
  const double uprev = d.b.u_prev[s];
  // QP model: optional velocity-form augmentation  A <- [A B; 0 1], B <- [B; 1], C <- C [I 0]
  KMPC_LANE_LOOP(e, nzq * nzq) {
    const int i = e / nzq, j = e - i * nzq;
    double v;
    if (i < nz) v = (j < nz) ? A[i * nz + j] : B[i];
    else v = (j == nz) ? 1.0 : 0.0;
    ws.A[e] = v;
  }
  KMPC_LANE_LOOP(e, nzq) {
    ws.B[e] = (e < nz) ? B[e] : 1.0;
    ws.z0[e] = (e < nz) ? d.b.z[s * nz + e] : uprev;
  }
  if (!identity) {
    KMPC_LANE_LOOP(e, ny * nzq) {
      const int i = e / nzq, j = e - i * nzq;
      const int row = (sh.out_mode == KMPC_OUT_C) ? i : c.out_row;
      ws.Cy[e] = (j < nz) ? C[row * nz + j] : 0.0;
    }
  }
  KMPC_LANE_LOOP(e, N) {
    double lo = c.lb, hi = c.ub;
    if (sh.du_aug && e == 0) {  // Tank_System.m:182-188: umin <= U0 + dU_1 <= umax
      lo = fmax(lo, c.u_lb - uprev);
      hi = fmin(hi, c.u_ub - uprev);
    }
    ws.lb[e] = lo;
    ws.ub[e] = hi;
  }
  KMPC_SYNCWARP();
  qp_build_warp<G, NZQ>(ws, nzq, ny, N, identity, c.q, c.rw, d.b.r + s * ny, 0, nullptr);
  // warm start: last step's optimal moves shifted by one (receding horizon; velocity form: the new
  // last move is "hold the input").  Uniform over the warp: one cold scenario makes its warp cold.
  bool warm = d.qp_x != nullptr;
  if (warm) warm = !warp_any(!isfinite(d.qp_x[s * N]));
  if (warm) {
    const double* xp = d.qp_x + s * N;
    KMPC_LANE_LOOP(i, N) ws.x[i] = (i + 1 < N) ? xp[i + 1] : (sh.du_aug ? 0.0 : xp[i]);
    KMPC_SYNCWARP();
  }
  const int st = qp_solve_warp<G, NMAX>(ws, N, c.max_iter, c.tol, warm, c.qp_cold != 2, G == 32 && c.qp_cold == 3);
  if (d.qp_x != nullptr && valid) {
    const bool ok = !(st & (KMPC_STATUS_NONFINITE | KMPC_STATUS_MAXITER));   // else: cold start next step
    KMPC_LANE_LOOP(i, N) d.qp_x[s * N + i] = ok ? ws.x[i] : NAN;
  }
  if (KMPC_LANE0 && valid) {
    const double move = ws.x[0];
    const double u = sh.du_aug ? uprev + move : move;
    const double* pp = (step < c.first_post_step ? d.b.params_pre : d.b.params_post) + s * 5;
    double p[5];
    for (int k = 0; k < 5; ++k) p[k] = pp[k];
    const double x1 = d.b.x[s * 2], x2 = d.b.x[s * 2 + 1];
    double o1, o2;
    plant_step_dev(c.plant_kind, c.rk4_variant, c.h, p, x1, x2, u, o1, o2);
    d.x_prev[s * 2] = x1;
    d.x_prev[s * 2 + 1] = x2;
    d.b.x[s * 2] = o1;
    d.b.x[s * 2 + 1] = o2;
    d.b.u_prev[s] = u;
    if (d.b.log_x && log_slot >= 0) {
      d.b.log_x[(log_slot * c.S + s) * 2] = o1;
      d.b.log_x[(log_slot * c.S + s) * 2 + 1] = o2;
    }
    if (d.b.log_u && log_slot >= 0) d.b.log_u[log_slot * c.S + s] = u;
    const int st2 = st | ((isfinite(o1) && isfinite(o2)) ? 0 : KMPC_STATUS_NONFINITE);   // plant left the reals
    if (d.b.status && st2) d.b.status[s] |= st2;
  }
}

// RLS with the sample (z, u) -> z_next; afterwards z <- z_next.  `first` = this is the restart
// update (duffing.py:927-930, 943-946): state is initialised to P = p0 I, bar_Q = q0 I, K_A = 0,
// bar_X = 0 instead of being read.
template <int G>
KMPC_DEV void loop_rls_scenario(const LoopDev& d, const int nz, int64_t s, bool valid, int first,
                                double* base) {
  const kmpc_loop_config& c = d.c;
  const int n = 2, nv = nz + 1;
  RlsWs ws = rls_ws_carve(base, nz, n);
  const bool upc = c.rls_flags & KMPC_RLS_UPDATE_C;
  if (first) {
    KMPC_LANE_LOOP(e, nz * nv) ws.KA[e] = 0.0;
    KMPC_LANE_LOOP(e, nv * nv) ws.P[e] = (e / nv == e % nv) ? c.p0 : 0.0;
    KMPC_LANE_LOOP(e, n * nz) ws.barX[e] = 0.0;
    KMPC_LANE_LOOP(e, nz * nz) ws.barQ[e] = (e / nz == e % nz) ? c.q0 : 0.0;
  } else {
    KMPC_COPY_G2S(ws.KA, d.b.KA + s * nz * nv, nz * nv);
    KMPC_COPY_G2S(ws.P, d.b.P + s * nv * nv, nv * nv);
    if (upc) {
      KMPC_COPY_G2S(ws.barX, d.b.barX + s * n * nz, n * nz);
      KMPC_COPY_G2S(ws.barQ, d.b.barQ + s * nz * nz, nz * nz);
    }
  }
  // the RLS state above is not touched by the QP / lift kernels of this step, so it was loaded
  // before waiting on them (programmatic dependent launch); the sample below is their output
  KMPC_PDL_WAIT();
  KMPC_LANE_LOOP(e, nz) {
    ws.v[e] = d.b.z[s * nz + e];
    ws.y[e] = d.z_next[s * nz + e];
  }
  if (KMPC_LANE0) ws.v[nz] = d.b.u_prev[s];
  KMPC_LANE_LOOP(e, n) ws.xc[e] = c.c_pairs_next ? d.b.x[s * n + e] : d.x_prev[s * n + e];
  KMPC_COPY_G2S_WAIT();
  KMPC_SYNCWARP();
  int flags = c.rls_flags;
  if (first && c.skip_first_barx) flags |= KMPC_RLS_SKIP_BARX;
  rls_update_warp<G>(ws, nz, n, c.lambda, flags);
  if (!valid) return;
  KMPC_LANE_LOOP(e, nz * nz) d.b.A[s * nz * nz + e] = ws.oA[e];
  KMPC_LANE_LOOP(e, nz) d.b.B[s * nz + e] = ws.oB[e];
  if (upc) KMPC_LANE_LOOP(e, n * nz) d.b.C[s * n * nz + e] = ws.oC[e];
  KMPC_LANE_LOOP(e, nz * nv) d.b.KA[s * nz * nv + e] = ws.KA[e];
  KMPC_LANE_LOOP(e, nv * nv) d.b.P[s * nv * nv + e] = ws.P[e];
  if (upc) {
    KMPC_LANE_LOOP(e, n * nz) d.b.barX[s * n * nz + e] = ws.barX[e];
    KMPC_LANE_LOOP(e, nz * nz) d.b.barQ[s * nz * nz + e] = ws.barQ[e];
  }
  KMPC_LANE_LOOP(e, nz) d.b.z[s * nz + e] = ws.y[e];
}

}  // namespace kmpc
