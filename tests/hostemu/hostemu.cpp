// hostemu.cpp -- TEST-ONLY host build of the warp-per-scenario device functions (percase.cuh
// compiled with KMPC_HOSTEMU: lane-strided loops run sequentially, warp primitives are no-ops).
// It lets `pytest -m "not gpu"` check the kernel *logic* (indexing, algorithm, status flags)
// against the oracle on a machine without a GPU.  It is built by tests/conftest.py into
// tests/hostemu/_build/, never shipped in the package and never loaded by the product path.
#define KMPC_HOSTEMU 1
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../koopman_online_updated_mpc_b200/csrc/loopbody.cuh"

using namespace kmpc;

extern "C" {

int emu_rls_update(double* KA, double* P, double* barX, double* barQ, const double* z,
                   const double* u, const double* y, const double* xc, double* A, double* B,
                   double* C, int64_t S, int nz, int n, double lambda, int flags) {
  const int nv = nz + 1;
  std::vector<double> buf(rls_ws_doubles(nz, n));
  for (int64_t s = 0; s < S; ++s) {
    RlsWs ws = rls_ws_carve(buf.data(), nz, n);
    memcpy(ws.KA, KA + s * nz * nv, sizeof(double) * nz * nv);
    memcpy(ws.P, P + s * nv * nv, sizeof(double) * nv * nv);
    if (flags & KMPC_RLS_UPDATE_C) {
      memcpy(ws.barX, barX + s * n * nz, sizeof(double) * n * nz);
      memcpy(ws.barQ, barQ + s * nz * nz, sizeof(double) * nz * nz);
      memcpy(ws.xc, xc + s * n, sizeof(double) * n);
    }
    memcpy(ws.v, z + s * nz, sizeof(double) * nz);
    memcpy(ws.y, y + s * nz, sizeof(double) * nz);
    ws.v[nz] = u[s];
    rls_update_warp<32>(ws, nz, n, lambda, flags);
    memcpy(A + s * nz * nz, ws.oA, sizeof(double) * nz * nz);
    memcpy(B + s * nz, ws.oB, sizeof(double) * nz);
    if ((flags & KMPC_RLS_UPDATE_C) && C) memcpy(C + s * n * nz, ws.oC, sizeof(double) * n * nz);
    memcpy(KA + s * nz * nv, ws.KA, sizeof(double) * nz * nv);
    memcpy(P + s * nv * nv, ws.P, sizeof(double) * nv * nv);
    if (flags & KMPC_RLS_UPDATE_C) {
      memcpy(barX + s * n * nz, ws.barX, sizeof(double) * n * nz);
      memcpy(barQ + s * nz * nz, ws.barQ, sizeof(double) * nz * nz);
    }
  }
  return 0;
}

int emu_qp_first_move(const double* A, const double* B, const double* Cy, const double* z0,
                      const double* r, const double* lb, const double* ub, const double* PN,
                      double q, double rw, int N, int ny, int nz, int64_t S, int flags, double* u0,
                      double* Ufull, int* status, int max_iter, double tol, double* Hout,
                      double* fout) {
  const bool identity = flags & KMPC_QP_CY_IDENTITY;
  const bool shared_model = flags & KMPC_QP_SHARED_MODEL;
  const bool r_full = flags & KMPC_QP_R_FULL;
  if (max_iter <= 0) max_iter = 10 * N + 20;
  if (!(tol > 0.0)) tol = 1e-10;
  std::vector<double> buf(qp_ws_doubles(nz, ny, N, identity) + 8);
  for (int64_t s = 0; s < S; ++s) {
    QpWs ws = qp_ws_carve(buf.data(), nz, ny, N, identity);
    const int64_t sm = shared_model ? 0 : s;
    memcpy(ws.A, A + sm * nz * nz, sizeof(double) * nz * nz);
    memcpy(ws.B, B + sm * nz, sizeof(double) * nz);
    memcpy(ws.z0, z0 + s * nz, sizeof(double) * nz);
    if (!identity) memcpy(ws.Cy, Cy + sm * ny * nz, sizeof(double) * ny * nz);
    memcpy(ws.lb, lb + s * N, sizeof(double) * N);
    memcpy(ws.ub, ub + s * N, sizeof(double) * N);
    const double* rs = r_full ? r + s * N * ny : r + s * ny;
    const double* pn = PN ? PN + sm * ny * ny : nullptr;
    qp_build_warp<32>(ws, nz, ny, N, identity, q, rw, rs, r_full ? ny : 0, pn);
    if (Hout)
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) Hout[(s * N + i) * N + j] = ws.H[i >= j ? tri(i, j) : tri(j, i)];
    if (fout) memcpy(fout + s * N, ws.f, sizeof(double) * N);
    const int st = qp_solve_warp<32>(ws, N, max_iter, tol);
    u0[s] = ws.x[0];
    if (status) status[s] = st;
    if (Ufull) memcpy(Ufull + s * N, ws.x, sizeof(double) * N);
  }
  return 0;
}

int emu_plant_step(const double* x, const double* u, const double* params, double* xnext,
                   int64_t S, int kind, int rk4_variant, double h) {
  for (int64_t s = 0; s < S; ++s)
    plant_step_dev(kind, rk4_variant, h, params + 5 * s, x[2 * s], x[2 * s + 1], u[s], xnext[2 * s],
                   xnext[2 * s + 1]);
  return 0;
}

int emu_rbf_lift(const double* x, const double* cx, double* z, int64_t S, int n, int nz, int variant) {
  for (int64_t s = 0; s < S; ++s)
    for (int c = 0; c < nz; ++c) z[s * nz + c] = rbf_thinplate(x + s * n, cx + c * n, n, variant);
  return 0;
}

// X = Bm inv(G); G (n x n) is destroyed
int emu_spd_right_solve(double* G, int n, double* Bm, int rows) {
  return spd_right_solve_warp<32>(G, n, Bm, rows);
}

// ---- fused closed loop, same schedule as kmpc_closed_loop_steps (closed_loop.cu): per step
// qp_plant -> lift -> rls.  The MLP lift is a plain host loop here (the CUDA encoder kernel is a
// block-cooperative GEMM that has no lane-emulation; it is checked on the GPU).
static void host_mlp(int n_layers, const int* dims, const double* const* W, const double* const* b,
                     const double* x, double* out) {
  std::vector<double> h(x, x + dims[0]), t;
  for (int l = 0; l < n_layers; ++l) {
    t.assign(dims[l + 1], 0.0);
    for (int o = 0; o < dims[l + 1]; ++o) {
      double s = b[l][o];
      for (int k = 0; k < dims[l]; ++k) s += W[l][(size_t)o * dims[l] + k] * h[k];
      t[o] = (l + 1 < n_layers && s < 0.0) ? 0.0 : s;
    }
    h.swap(t);
  }
  for (size_t i = 0; i < h.size(); ++i) out[i] = h[i];
}

static void host_lift(const kmpc_loop_config& c, const kmpc_loop_buffers& b, int n_layers,
                      const int* dims, const double* const* W, const double* const* bias,
                      const double* x, double* z) {
  if (c.lift_kind == KMPC_LIFTKIND_RBF) {
    for (int k = 0; k < c.nz; ++k) z[k] = rbf_thinplate(x, b.cx + k * c.n, c.n, c.lift_mode);
    return;
  }
  const int nzo = dims[n_layers];
  std::vector<double> t(nzo), t0(nzo), zero(dims[0], 0.0);
  host_mlp(n_layers, dims, W, bias, x, t.data());
  if (c.lift_mode == KMPC_LIFT_RAW) {
    for (int k = 0; k < nzo; ++k) z[k] = t[k];
    return;
  }
  host_mlp(n_layers, dims, W, bias, zero.data(), t0.data());
  int off = 0;
  if (c.lift_mode == KMPC_LIFT_STACK) {
    for (int k = 0; k < dims[0]; ++k) z[k] = x[k];
    off = dims[0];
  }
  for (int k = 0; k < nzo; ++k) z[off + k] = t[k] - t0[k];
}

int emu_closed_loop(const kmpc_loop_config* cfg, const kmpc_loop_buffers* buf, int T,
                    int rls_started, int64_t start_step, int n_layers, const int* dims,
                    const double* const* W, const double* const* bias, double* qp_x) {
  LoopDev d;
  d.c = *cfg;
  d.b = *buf;
  if (d.c.max_iter <= 0) d.c.max_iter = 10 * d.c.N + 20;
  if (!(d.c.tol > 0.0)) d.c.tol = 1e-10;
  const kmpc_loop_config& c = d.c;
  std::vector<double> znext((size_t)c.S * c.nz), xprev((size_t)c.S * c.n);
  d.z_next = znext.data();
  d.x_prev = xprev.data();
  d.wset = nullptr;
  d.qp_x = qp_x;   // (S, N) caller-owned warm-start move sequences (NaN = none), nullable (always cold)
  const bool identity = c.out_mode == KMPC_OUT_IDENTITY;
  std::vector<double> qws(qp_ws_doubles(loop_nzq(c), loop_ny(c), c.N, identity) + 8);
  std::vector<double> rws(rls_ws_doubles(c.nz, c.n) + 8);
  int64_t step = start_step;
  for (int t = 0; t < T; ++t, ++step) {
    const int64_t slot = step < d.b.log_capacity ? step : -1;
    for (int64_t s = 0; s < c.S; ++s) loop_qp_plant_scenario<32>(d, loop_shape(c), s, true, step, slot, qws.data());
    double* zdst = c.update ? d.z_next : d.b.z;
    for (int64_t s = 0; s < c.S; ++s)
      host_lift(c, d.b, n_layers, dims, W, bias, d.b.x + s * c.n, zdst + s * c.nz);
    if (c.update) {
      for (int64_t s = 0; s < c.S; ++s) loop_rls_scenario<32>(d, c.nz, s, true, rls_started ? 0 : 1, rws.data());
      rls_started = 1;
    }
  }
  return 0;
}

}  // extern "C"
