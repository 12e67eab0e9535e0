// fused.cu -- the persistent fused closed-loop kernel (nz = 8, horizon 10: duffing.py,
// vanderpol.py, duffing_RBF.py shapes).
//
// ONE launch runs T closed-loop steps of every scenario (duffing.py:823-992 loop body):
//     z = lift(x) -> condensed box-QP -> u -> x+ = plant(x, u) -> y = lift(x+) -> RLS(z,u,y) -> A,B,C
// A CTA owns a tile of 32 scenarios for all T steps; 8 lanes own one scenario.  Everything a
// scenario carries from step to step -- the model A, B, C, the lift z, the plant state and the RLS
// state K_A, P, bar_X, bar_Q (233 doubles) -- lives in the REGISTERS of its 8 lanes, distributed by
// rows (lane i holds row i), and touches HBM once per launch instead of once per step.  The
// encoder weights (175 KB) are brought into shared memory once per CTA by the TMA engine and the
// 2-100-100-100-8 MLP of a step runs on the fp64 tensor path (encoder.cuh: lift_unit).
//
// A QUARTER (lift unit) = 2 warps = 8 scenarios = one DMMA m-tile.  It owns a region of shared memory:
// two ping-pong activation buffers aliased by its scenarios' scratch (184 doubles each; the scratch
// is dead while the unit lifts), the layer-0 input block and the lift outputs.  Quarters
// synchronise only internally (named barrier, 64 threads): there is no CTA-wide barrier in the
// step loop and the four quarters drift freely.
//
// Step schedule of a quarter (warp w = scenarios 4w..4w+3):
//   QP build   lanes cooperate: Krylov chains VZ[t] = A^(t+1) z, VB[t] = A^t B (lane i = component
//              i, vectors exchanged through shared memory), then H = q G'G + rw I and
//              f = 2 q G'(F z - r) from lane-local partial sums reduced across the 8 lanes;
//   QP solve   8 lanes (QpCoop): exact primal-dual active-set solve (same algorithm as percase.cuh
//              qp_solve_warp / oracle solve_box_qp_exact), compile-time horizon, warm-started from
//              the previous step's working set, right-looking Cholesky with replicated pivots;
//   plant      lane 0: RK4 / tank map, logs;
//   lift       quarter: lift_unit<2> on the 8 new states (tensor path) or 8 RBFs per scenario;
//   RLS        lanes cooperate: Sherman-Morrison on P and bar_Q, K_A += y v', [A B] = K_A P,
//              C = bar_X bar_Q (duffing.py:927-953), formula order of the reference.
//
// Shapes outside this kernel's specialisation (Tank: nz = 10 + du augmentation, N = 20; N = 50)
// run on the generic three-kernel path of closed_loop.cu.
#include "encoder.cuh"
#include "loopbody.cuh"

namespace kmpc {

__device__ __forceinline__ void quarter_barrier(int id) { group_barrier<64>(id); }

constexpr int FG = 8;            // lanes per scenario
constexpr int FNZ = 8;           // lifted dimension
constexpr int FN = 10;           // horizon
constexpr int FNV = FNZ + 1;
constexpr int kScr = 184;        // scratch doubles per scenario (== 8 mod 16: the two scenarios of a
                                 // half-warp land on disjoint banks)
constexpr unsigned kSetMask = (1u << FN) - 1u, kKeepBit = 1u << 31;   // wset word 0: working set | warm-start policy
constexpr int kRedPitch = 9;     // pitch of the 8 x 8 reduction buffer (conflict-free column reads)
// scratch layout during the QP build and the RLS
constexpr int oRED = 0;          // [8][kRedPitch] cross-lane reduction buffer
constexpr int oEX = 72;          // 40 doubles: vector exchange (Krylov pairs / g,e / RLS gathers)
constexpr int oHF = 112;         // 72 doubles: 2H packed lower triangle (55), f (10), pad
// scratch layout during the QP solve (aliases RED/EX; HF stays live)
constexpr int oL = 0;            // 60: strictly-lower factor + forward-substituted rhs, column major
constexpr int oXS = 60;          // 10: current iterate x
constexpr int oGS = 70;          // 10: gradient 2 H x + f
// unit region (one per quarter = 8 scenarios): [ bufA | bufB ] ping-pong activations of the unit's
// lift, aliased by the unit's scratch (8 x kScr doubles from the start: the scratch is dead while the
// unit lifts); the tail beyond the scratch holds the layer-0 input block (4 x 8) and the lift outputs
// (8 x 8); the split-K partials of the last layer use the ping-pong buffer that layer does not read
constexpr int kIn0 = 4 * kUnitRows, kYout = kUnitRows * FNZ;

// The identity-output build emits the 55 + 10 reduced values diagonal by diagonal (running sums
// along a diagonal); this table maps emission number -> position in HF (packed lower triangle
// tri(a, b) for H, 55 + a for f, identity for the 7 padding slots).
__constant__ unsigned char c_hf_tab[72] = {
    54, 44, 35, 27, 20, 14, 9, 5, 2, 0, 53, 43, 34, 26, 19, 13, 8, 4, 1, 52, 42, 33, 25, 18, 12, 7, 3, 51,
    41, 32, 24, 17, 11, 6, 50, 40, 31, 23, 16, 10, 49, 39, 30, 22, 15, 48, 38, 29, 21, 47, 37, 28, 46, 36,
    45, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64, 65, 66, 67, 68, 69, 70, 71};

// first element of column j of the factor: rows j+1 .. 10 (row 10 = the right-hand side carried
// through the elimination), padded to an even length
__host__ __device__ constexpr int lcol(int j) {
  int off = 0;
  for (int k = 0; k < j; ++k) off += ((FN - k) + 1) & ~1;
  return off;
}

struct FusedSmem {  // offsets in doubles from the start of dynamic shared memory
  int region;       // doubles per unit region (4 regions from offset 0)
  int actbuf;       // doubles per ping-pong buffer
  int in0, yout;    // offsets of the layer-0 input block / lift outputs inside a region
  int wsm, bars, total_bytes;
};
inline FusedSmem fused_smem_layout(const EncParams* p) {
  FusedSmem L;
  L.actbuf = p ? p->actw * kActPitch : 0;
  const int scratch = kUnitRows * kScr;
  L.in0 = scratch;
  L.yout = scratch + kIn0;
  L.region = scratch + kIn0 + kYout;
  if (2 * L.actbuf > L.region) L.region = 2 * L.actbuf;
  L.region = (L.region + 1) & ~1;
  L.wsm = 4 * L.region;
  L.bars = L.wsm + (p ? p->total_w : 0);
  L.bars = (L.bars + 1) & ~1;
  L.total_bytes = (L.bars + KMPC_MAX_LAYERS) * 8;
  return L;
}

// sum over the 8 lanes of a scenario of value number `l`: lane l of every group receives the
// reduced value whose partials the lanes passed in ch[l].  Partials are added in lane order.
__device__ __forceinline__ double chunk_reduce(double* red, int l, const double (&ch)[8]) {
#pragma unroll
  for (int v = 0; v < 8; ++v) red[v * kRedPitch + l] = ch[v];
  __syncwarp();
  const double* r = red + l * kRedPitch;
  const double s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  __syncwarp();
  return s;
}

// every lane of the scenario receives all 8 lane values (in lane order) through `ex`
__device__ __forceinline__ void group_gather(double* ex, int l, double mine, double (&all)[8]) {
  ex[l] = mine;
  __syncwarp();
  const double2* e2 = reinterpret_cast<const double2*>(ex);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double2 t = e2[j];
    all[2 * j] = t.x;
    all[2 * j + 1] = t.y;
  }
  __syncwarp();
}

// ---------------------------------------------------------------- cooperative exact box-QP ---
// min x'Hx + f'x, lo <= x <= hi for one scenario, executed by its 8 lanes; the 4 scenarios of a
// warp run in lock step (every loop is warp-uniform; a scenario that has converged keeps executing
// with its state frozen).  Same algorithm as percase.cuh qp_solve_warp / oracle
// solve_box_qp_exact (primal-dual active-set sweeps, then the monotone primal method), started
// from a warm working set.
//   factor_solve: Cholesky of the free block of 2H with the right-hand side carried along, done by
//   every lane redundantly in registers (no exchange, no synchronisation), which leaves the step p
//   replicated in registers.
struct QpCoop {
  const double* HF;   // packed lower 2H (55) | f (10)
  double* sc;         // Lc (60) | xs (10) | gs (10)
  int l;              // lane within the scenario
  int status;
  int iters;          // active-set iterations of the last run()
  unsigned wlo, whi;  // working set: bit i set = variable i at its lower / upper bound

  // p <- (2H)_FF^-1 rhs_F (0 on masked variables), rhs = -g on free variables; replicated result.
  // Every lane factors the whole (masked) 10 x 10 block in its own registers, left-looking: no
  // exchange and no synchronisation inside the factorisation, and the 9 - j row chains of a column are
  // independent (the cooperative right-looking form needed one shared-memory round trip per column;
  // the redundant lanes cost nothing in a SIMT warp).  The sequence of roundings per entry is the one
  // of the right-looking form: s = H_rj, then s = fma(-L_rk, L_jk, s) for k = 0 .. j-1, then s * inv_j.
  template <bool MASKED, int J>
  __device__ __forceinline__ void factor_column(unsigned masked, const double* gs, double* Lc,
                                                double (&L)[(FN * (FN - 1)) / 2], double (&invd)[FN],
                                                double (&p)[FN], int& st) {
    constexpr int bj = (J * (J - 1)) / 2;
    const bool mj = MASKED && ((masked >> J) & 1u);
    const double hjj = HF[(J * (J + 1)) / 2 + J];
    double d = mj ? 1.0 : hjj;
    double y = mj ? 0.0 : -gs[J];
#pragma unroll
    for (int k = 0; k < J; ++k) {
      d = fma(-L[bj + k], L[bj + k], d);
      y = fma(-p[k], L[bj + k], y);
    }
    const double floor_j = mj ? 0.0 : kPivotFloor * hjj;
    if (!(d > floor_j)) {  // numerically semi-definite: regularise and flag (oracle/mpc.py PIVOT_FLOOR)
      st |= KMPC_STATUS_PIVOT;
      d = floor_j;
    }
    const double inv = rsqrt_pos(d);
    invd[J] = inv;
    p[J] = y * inv;   // y_J (forward substitution rides along)
#pragma unroll
    for (int r = J + 1; r < FN; ++r) {
      const int br = (r * (r - 1)) / 2;
      double s = (mj || (MASKED && ((masked >> r) & 1u))) ? 0.0 : HF[(r * (r + 1)) / 2 + J];
#pragma unroll
      for (int k = 0; k < J; ++k) s = fma(-L[br + k], L[bj + k], s);
      const double lrj = s * inv;
      L[br + J] = lrj;
      if (l == 0) Lc[br + J] = lrj;   // parked for the back substitution: row r leaves the registers after column r
    }
  }

  // MASKED = false: empty working set (no selects at all), right-hand side -f straight from HF
  template <bool MASKED>
  __device__ __forceinline__ int factor_solve(unsigned masked, double (&p)[FN]) {
    static_assert(FN == 10, "factor_solve instantiates the ten columns by hand");
    const double* gs = MASKED ? sc + oGS : HF + 55;
    double* Lc = sc;
    double L[(FN * (FN - 1)) / 2], invd[FN];   // strictly lower triangle, row r at r (r - 1) / 2
    int st = 0;
    factor_column<MASKED, 0>(masked, gs, Lc, L, invd, p, st);
    factor_column<MASKED, 1>(masked, gs, Lc, L, invd, p, st);
    factor_column<MASKED, 2>(masked, gs, Lc, L, invd, p, st);
    factor_column<MASKED, 3>(masked, gs, Lc, L, invd, p, st);
    factor_column<MASKED, 4>(masked, gs, Lc, L, invd, p, st);
    factor_column<MASKED, 5>(masked, gs, Lc, L, invd, p, st);
    factor_column<MASKED, 6>(masked, gs, Lc, L, invd, p, st);
    factor_column<MASKED, 7>(masked, gs, Lc, L, invd, p, st);
    factor_column<MASKED, 8>(masked, gs, Lc, L, invd, p, st);
    factor_column<MASKED, 9>(masked, gs, Lc, L, invd, p, st);
    __syncwarp();
    // back substitution L' x = y
#pragma unroll
    for (int jj = 0; jj < FN; ++jj) {
      const int j = FN - 1 - jj;
      double s0 = p[j], s1 = 0.0;
#pragma unroll
      for (int r = j + 1; r < FN; ++r) {
        const double lrj = Lc[((r * (r - 1)) >> 1) + j];
        if ((r - j) & 1) s0 = fma(-lrj, p[r], s0);
        else s1 = fma(-lrj, p[r], s1);
      }
      p[j] = (s0 + s1) * invd[j];
    }
    __syncwarp();   // Lc is rewritten by the next factorisation
    return st;
  }

  // gs = 2 H xs + f: lane i computes rows i (and i + 8), then the group exchanges through gs
  __device__ __forceinline__ void gradient() {
    const double* xs = sc + oXS;
    double* gs = sc + oGS;
    double xv[FN];
#pragma unroll
    for (int j = 0; j < FN; ++j) xv[j] = xs[j];
#pragma unroll
    for (int slot = 0; slot < 2; ++slot) {
      const int r = l + 8 * slot;
      if (r < FN) {
        const int br = (r * (r + 1)) >> 1;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int j = 0; j < FN; j += 2) {
          s0 = fma(HF[j <= r ? br + j : tri(j, 0) + r], xv[j], s0);
          s1 = fma(HF[j + 1 <= r ? br + j + 1 : tri(j + 1, 0) + r], xv[j + 1], s1);
        }
        gs[r] = (s0 + s1) + HF[55 + r];
      }
    }
    __syncwarp();
  }

  // Returns the first move; leaves the final working set in (wlo, whi).  Every lane of the
  // scenario returns the same value.
  __device__ __forceinline__ double run(double lo, double hi, int max_iter, double tol) {
    status = 0;
    iters = 0;
    double* xs = sc + oXS;
    double* gs = sc + oGS;
    double fmaxabs = 0.0;
#pragma unroll
    for (int i = 0; i < FN; ++i) fmaxabs = fmax(fmaxabs, fabs(HF[55 + i]));
    const double mtol = tol * fmax(1.0, fmaxabs);
    const double x0 = fmin(fmax(0.0, lo), hi);
    const bool cold = (wlo | whi) == 0u && x0 == 0.0;
    if (l == 0) {
#pragma unroll
      for (int i = 0; i < FN; ++i) {
        xs[i] = ((wlo >> i) & 1u) ? lo : (((whi >> i) & 1u) ? hi : x0);
        gs[i] = HF[55 + i];    // gradient at x = 0 (cold start)
      }
    }
    __syncwarp();
    // warm start: gradient at the start point (for a cold scenario x = 0 and this reproduces g = f)
    if (__any_sync(0xffffffffu, !cold)) gradient();
    bool done = false;
    for (int it = 0; it < max_iter; ++it) {
      if (!__any_sync(0xffffffffu, !done)) break;
      if (!done) ++iters;
      const bool pdas = it < kPdasIters;
      const unsigned masked = wlo | whi;
      double p[FN];
      const int st = factor_solve<true>(masked, p);
      if (!done) status |= st;
      double alpha = 1.0;
      int block = -1;
      if (!pdas) {  // ratio test; lowest index wins ties
#pragma unroll
        for (int i = 0; i < FN; ++i) {
          if (!((masked >> i) & 1u)) {
            const double pi = p[i], xi = xs[i];
            double a = 2.0;
            if (pi > 0.0 && xi + pi > hi) a = (hi - xi) / pi;
            else if (pi < 0.0 && xi + pi < lo) a = (lo - xi) / pi;
            if (a < alpha) {
              alpha = a;
              block = i;
            }
          }
        }
      }
      unsigned nlo = wlo, nhi = whi;
      double xn[FN];
#pragma unroll
      for (int i = 0; i < FN; ++i) {
        double xi = fma(alpha, p[i], xs[i]);
        if (i == block) {
          if (p[i] > 0.0) {
            xi = hi;
            nhi |= 1u << i;
          } else {
            xi = lo;
            nlo |= 1u << i;
          }
        }
        xn[i] = xi;
      }
      __syncwarp();   // every lane has read xs
      if (l == 0 && !done) {
#pragma unroll
        for (int i = 0; i < FN; ++i) xs[i] = xn[i];
      }
      __syncwarp();
      bool finished = false;
      if (pdas) {
        // multipliers (gradient at the unclipped face minimiser) are needed only when something
        // is bound; the free components of the gradient vanish there
        if (__any_sync(0xffffffffu, !done && masked != 0u)) gradient();
        bool changed = false, clipped = false;
#pragma unroll
        for (int i = 0; i < FN; ++i) {
          const unsigned bit = 1u << i;
          if (masked & bit) {       // release every bound whose multiplier has the wrong sign
            const double lamb = (nlo & bit) ? gs[i] : -gs[i];
            if (lamb < -mtol) {
              nlo &= ~bit;
              nhi &= ~bit;
              changed = true;
            }
          } else if (xn[i] < lo) {  // clip every violated bound into the working set
            xn[i] = lo;
            nlo |= bit;
            clipped = true;
          } else if (xn[i] > hi) {
            xn[i] = hi;
            nhi |= bit;
            clipped = true;
          }
        }
        finished = !changed && !clipped;
        __syncwarp();
        if (l == 0 && !done && clipped) {
#pragma unroll
          for (int i = 0; i < FN; ++i) xs[i] = xn[i];
        }
        __syncwarp();
        // next iteration starts from the gradient at the clipped point
        if (__any_sync(0xffffffffu, !done && !finished && (clipped || masked == 0u))) gradient();
      } else {
        gradient();
        double worst = INFINITY;
        int widx = -1;
#pragma unroll
        for (int i = 0; i < FN; ++i) {
          const unsigned bit = 1u << i;
          if ((nlo | nhi) & bit) {
            const double lamb = (nlo & bit) ? gs[i] : -gs[i];
            if (lamb < worst) {
              worst = lamb;
              widx = i;
            }
          }
        }
        if (block < 0) {
          if (widx < 0 || worst >= -mtol) {
            finished = true;
          } else {
            nlo &= ~(1u << widx);
            nhi &= ~(1u << widx);
          }
        }
      }
      if (!done) {   // converged scenarios keep their state frozen
        wlo = nlo;
        whi = nhi;
        done = finished;
      }
    }
    if (!done) status |= KMPC_STATUS_MAXITER;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < FN; ++i) bad |= !isfinite(xs[i]);
    if (bad) {
      status |= KMPC_STATUS_NONFINITE;
      wlo = whi = 0u;
    }
    const double u0 = xs[0];
    __syncwarp();
    return u0;
  }
};

// ---------------------------------------------------------------- the kernel -----------------
// OUT: KMPC_OUT_IDENTITY (y = z, vanderpol.py) or KMPC_OUT_C (y = C z, duffing.py);
// UPDATE: online RLS; MLP: theta_E lift (else thin-plate RBF); TIMED: per-phase clock64 sums.
struct FusedArgs {
  LoopDev d;
  EncParams p;        // MLP only
  FusedSmem sm;
  int64_t step0;      // closed-loop index of the first step of this launch
  int T;              // steps in this launch
  int first;          // 1: the first step restarts the RLS state (P = p0 I, bar_Q = q0 I)
  int64_t num_tiles;
  long long* timing;  // TIMED: [gridDim.x][4] cycles in QP+plant, lift, RLS, total
#ifdef KMPC_PROFILING
  int dbg_skip;       // profiling builds only (-DKMPC_PROFILING, env KMPC_FUSED_SKIP): bit 0 skip QP, bit 1 skip
                      // lift, bit 2 skip RLS, bit 3 log the active-set iteration count in log_u instead of u
#endif
};
// release builds have no phase-skip knob at all: the expression folds to 0
#ifdef KMPC_PROFILING
#define KMPC_DBG_SKIP(a) ((a).dbg_skip)
#else
#define KMPC_DBG_SKIP(a) 0
#endif

template <int OUT, bool UPDATE, bool MLP, bool TIMED>
__global__ void __launch_bounds__(kMmaThreads, 1) fused_loop_kernel(const __grid_constant__ FusedArgs a) {
  extern __shared__ __align__(16) double smem[];
  const kmpc_loop_config& c = a.d.c;
  const kmpc_loop_buffers& b = a.d.b;
  double* wsm = smem + a.sm.wsm;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.sm.bars);
  const int tid = threadIdx.x;
  const int sc = tid >> 3, l = tid & 7;
  // quarter = 2 warps = 8 scenarios = one lift unit: its own region (ping-pong activations aliasing
  // its scenarios' scratch, layer-0 block, lift outputs) and named barrier; quarters never wait for
  // each other inside the step loop.  Quarters q and q + 2 share two schedulers: the warp that owns
  // the odd 13th n-tile of a 100-wide layer alternates between them (13 tile columns per scheduler)
  const int quarter = tid >> 6, wl = (tid >> 5) & 1, qbar = 1 + quarter;
  double* region = smem + quarter * a.sm.region;
  double* scr = region + (sc & 7) * kScr;
  double* in0 = region + a.sm.in0;
  double* yout = region + a.sm.yout;
  double* red = scr + oRED;
  double* ex = scr + oEX;
  double* HF = scr + oHF;
  if (MLP) {
    if (tid == 0) encoder_weights_init_barriers(a.p, bars);
    __syncthreads();
    if (tid == 0) encoder_weights_issue(a.p, wsm, bars);
  }
  if (MLP) encoder_weights_wait_all(a.p, bars);   // overlaps nothing worth keeping: the first lift is a QP away
  const bool upc = (c.rls_flags & KMPC_RLS_UPDATE_C) != 0;
  const double lam = c.lambda;
  long long tq = 0, tl = 0, tr = 0, t_begin = 0;
  if (TIMED) t_begin = clock64();

  for (int64_t tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    int64_t s = tile * kTileS + sc;
    const bool valid = s < c.S;
    if (!valid) s = c.S - 1;
    const int64_t sm = c.shared_model ? 0 : s;
    // ---- state -> registers (lane l = row l) ----
    // A, B: frozen model -> registers for the whole launch; online update -> they are produced by
    // the RLS at the end of a step and consumed by the Krylov chains at the start of the next one,
    // so they wait in the (then idle) HF part of the scratch, transposed (lane-contiguous)
    double Ar[FNZ], Bl, Cc0 = 0.0, Cc1 = 0.0;
#pragma unroll
    for (int j = 0; j < FNZ; ++j) Ar[j] = b.A[sm * 64 + l * 8 + j];
    Bl = b.B[sm * 8 + l];
    if (UPDATE) {
#pragma unroll
      for (int j = 0; j < FNZ; ++j) HF[j * 8 + l] = Ar[j];
      HF[64 + l] = Bl;
    }
    if (OUT == KMPC_OUT_C || UPDATE) {
      Cc0 = b.C[sm * 16 + l];
      Cc1 = b.C[sm * 16 + 8 + l];
    }
    double zl = b.z[s * 8 + l];
    double x1 = b.x[s * 2], x2 = b.x[s * 2 + 1];
    double uprev = b.u_prev[s];
    double rl = 0.0, r0 = 0.0, r1 = 0.0;
    if (OUT == KMPC_OUT_IDENTITY) {
      rl = b.r[s * 8 + l];
    } else {
      r0 = b.r[s * 2];
      r1 = b.r[s * 2 + 1];
    }
    double KAr[FNV], Pr[FNV], P8l = 0.0, P88 = 0.0, Qr[FNZ], Xc0 = 0.0, Xc1 = 0.0;
    if (UPDATE) {
      if (a.first) {  // duffing.py:927-930, 943-946
#pragma unroll
        for (int j = 0; j < FNV; ++j) {
          KAr[j] = 0.0;
          Pr[j] = (j == l) ? c.p0 : 0.0;
        }
        P88 = c.p0;
#pragma unroll
        for (int j = 0; j < FNZ; ++j) Qr[j] = (j == l) ? c.q0 : 0.0;
      } else {
#pragma unroll
        for (int j = 0; j < FNV; ++j) {
          KAr[j] = b.KA[s * 72 + l * 9 + j];
          Pr[j] = b.P[s * 81 + l * 9 + j];
        }
        P8l = b.P[s * 81 + 72 + l];
        P88 = b.P[s * 81 + 80];
        if (upc) {
#pragma unroll
          for (int j = 0; j < FNZ; ++j) Qr[j] = b.barQ[s * 64 + l * 8 + j];
          Xc0 = b.barX[s * 16 + l];
          Xc1 = b.barX[s * 16 + 8 + l];
        } else {
#pragma unroll
          for (int j = 0; j < FNZ; ++j) Qr[j] = 0.0;
        }
      }
    }
    int status = 0;
    unsigned wlo = 0u, whi = 0u;   // optimal working set of the previous step (lane 0)
    if (a.d.wset) {
      const uint2 w2 = reinterpret_cast<const uint2*>(a.d.wset)[s];
      wlo = w2.x;
      whi = w2.y;
    }

    for (int t = 0; t < a.T; ++t) {
      const int64_t step = a.step0 + t;
      long long c0 = 0;
      if (TIMED) c0 = clock64();
      double x1n = x1, x2n = x2, unew = uprev;
      if (!(KMPC_DBG_SKIP(a) & 1)) {
      // ================= QP build: Krylov chains =================
      double VZ[FN], VB[FN];
      if (UPDATE) {
#pragma unroll
        for (int j = 0; j < FNZ; ++j) Ar[j] = HF[j * 8 + l];
        Bl = HF[64 + l];
      }
      VB[0] = Bl;
      ex[l] = zl;
      ex[8 + l] = Bl;
      __syncwarp();
#pragma unroll
      for (int k = 0; k < FN; ++k) {
        const double2* src = reinterpret_cast<const double2*>(ex + (k & 1) * 16);
        double zv[8], bv[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double2 tz = src[j];
          zv[2 * j] = tz.x;
          zv[2 * j + 1] = tz.y;
        }
        double sz = 0.0;
#pragma unroll
        for (int j = 0; j < FNZ; ++j) sz = fma(Ar[j], zv[j], sz);
        VZ[k] = sz;
        if (k + 1 < FN) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double2 tb = src[4 + j];
            bv[2 * j] = tb.x;
            bv[2 * j + 1] = tb.y;
          }
          double sb = 0.0;
#pragma unroll
          for (int j = 0; j < FNZ; ++j) sb = fma(Ar[j], bv[j], sb);
          VB[k + 1] = sb;
          double* dst = ex + ((k + 1) & 1) * 16;
          dst[l] = sz;
          dst[8 + l] = sb;
          __syncwarp();
        }
      }
      __syncwarp();
      // ================= QP build: H and f =================
      if (OUT == KMPC_OUT_IDENTITY) {
        // g[t] = VB[t], e[t] = VZ[t] - r: lane l holds component l; partial sums over its own
        // component, reduced over the 8 lanes in chunks of 8 values (emission order = HF order)
        double ch[8];
        int n = 0;
#pragma unroll
        for (int dd = 0; dd < FN; ++dd) {
          double run = 0.0;
#pragma unroll
          for (int k = 0; k + dd < FN; ++k) {
            run = fma(VB[k + dd], VB[k], run);
            ch[n & 7] = (2.0 * c.q) * run;
            ++n;
            if ((n & 7) == 0) {
              const double sum = chunk_reduce(red, l, ch);
              const int idx = n - 8 + l;
              HF[c_hf_tab[idx]] = sum + (idx < FN ? 2.0 * c.rw : 0.0);
            }
          }
        }
#pragma unroll
        for (int aa = 0; aa < FN; ++aa) {
          double sacc = 0.0;
#pragma unroll
          for (int k = aa; k < FN; ++k) sacc = fma(VB[k - aa], VZ[k] - rl, sacc);
          ch[n & 7] = 2.0 * c.q * sacc;
          ++n;
          if ((n & 7) == 0) {
            const double sum = chunk_reduce(red, l, ch);
            HF[c_hf_tab[n - 8 + l]] = sum;
          }
        }
        // n == 65: one value left in ch[0]
#pragma unroll
        for (int v = 1; v < 8; ++v) ch[v] = 0.0;
        {
          const double sum = chunk_reduce(red, l, ch);
          HF[64 + l] = sum;
        }
      } else {
        // g[t] = C VB[t], e[t] = C VZ[t] - r (ny = 2): 40 values reduced over the lanes into
        // ex[4 t + {0,1}] = g[t], ex[4 t + {2,3}] = e[t]
        double ge[5];
#pragma unroll
        for (int q5 = 0; q5 < 5; ++q5) {
          double ch[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k = 2 * q5 + h;
            ch[4 * h + 0] = Cc0 * VB[k];
            ch[4 * h + 1] = Cc1 * VB[k];
            ch[4 * h + 2] = Cc0 * VZ[k];
            ch[4 * h + 3] = Cc1 * VZ[k];
          }
          ge[q5] = chunk_reduce(red, l, ch);
        }
        // lane l received value l of every chunk: (l & 3) >= 2 are e entries
        const double roff = ((l & 3) == 2) ? r0 : (((l & 3) == 3) ? r1 : 0.0);
#pragma unroll
        for (int q5 = 0; q5 < 5; ++q5) ex[8 * q5 + l] = ge[q5] - roff;
        __syncwarp();
        // diagonals of H: pair-task pt = d and 9 - d (11 dot products); f likewise
        for (int task = l; task < 10; task += FG) {
          const bool isf = task >= 5;
          const int d0 = isf ? task - 5 : task;
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            const int dd = half ? (FN - 1 - d0) : d0;
            if (!isf) {
              double run = 0.0;
              for (int k = 0; k + dd < FN; ++k) {
                const double2 g1 = *reinterpret_cast<const double2*>(ex + 4 * (k + dd));
                const double2 g0 = *reinterpret_cast<const double2*>(ex + 4 * k);
                run = fma(g1.x, g0.x, run);
                run = fma(g1.y, g0.y, run);
                const int ra = FN - 1 - k;   // H[ra][ra - dd]
                HF[((ra * (ra + 1)) >> 1) + ra - dd] = 2.0 * (c.q * run + (dd == 0 ? c.rw : 0.0));
              }
            } else {
              double sacc = 0.0;  // f[a], a = dd
              for (int k = dd; k < FN; ++k) {
                const double2 g0 = *reinterpret_cast<const double2*>(ex + 4 * (k - dd));
                const double2 e1 = *reinterpret_cast<const double2*>(ex + 4 * k + 2);
                sacc = fma(g0.x, e1.x, sacc);
                sacc = fma(g0.y, e1.y, sacc);
              }
              HF[55 + dd] = 2.0 * c.q * sacc;
            }
          }
        }
      }
      __syncwarp();
      // ================= QP solve (8 lanes) + plant (lane 0) =================
      {
        QpCoop qp;
        qp.HF = HF;
        qp.sc = scr + oL;
        qp.l = l;
        // warm start: last step's optimal working set, either shifted by one move (receding horizon:
        // right when the plan is being followed) or as it is (right when the pattern is stationary in
        // the horizon frame, e.g. "first move free, the rest saturated" while the restarted model is
        // still poor: the shifted guess then costs 2-4 sweeps EVERY step).  The guess that matched the
        // optimum of the last step is used for the next one (kKeepBit of wlo, carried in wset).
        const unsigned plo = wlo & kSetMask, phi = whi & kSetMask;
        const unsigned slo = (plo >> 1) | (plo & (1u << (FN - 1))), shi = (phi >> 1) | (phi & (1u << (FN - 1)));
        const bool keep = (wlo & kKeepBit) != 0u;
        qp.wlo = keep ? plo : slo;
        qp.whi = keep ? phi : shi;
        unew = qp.run(c.lb, c.ub, c.max_iter, c.tol);
        status |= qp.status;
        const bool as_kept = qp.wlo == plo && qp.whi == phi, as_shifted = qp.wlo == slo && qp.whi == shi;
        const bool keep_next = (as_kept != as_shifted) ? as_kept : keep;
        wlo = qp.wlo | (keep_next ? kKeepBit : 0u);
        whi = qp.whi;
        if (l == 0) {
          const double* pp = (step < c.first_post_step ? b.params_pre : b.params_post) + s * 5;
          double prm[5];
#pragma unroll
          for (int k = 0; k < 5; ++k) prm[k] = __ldg(pp + k);
          plant_step_dev(c.plant_kind, c.rk4_variant, c.h, prm, x1, x2, unew, x1n, x2n);
          if (!isfinite(x1n) || !isfinite(x2n)) status |= KMPC_STATUS_NONFINITE;   // plant left the reals
          if (valid) {
            const int64_t slot = (step < b.log_capacity) ? step : -1;
            if (b.log_x && slot >= 0) {
              b.log_x[(slot * c.S + s) * 2] = x1n;
              b.log_x[(slot * c.S + s) * 2 + 1] = x2n;
            }
            if (b.log_u && slot >= 0) b.log_u[slot * c.S + s] = (KMPC_DBG_SKIP(a) & 8) ? (double)qp.iters : unew;
          }
        }
      }
      __syncwarp();
      x1n = __shfl_sync(0xffffffffu, x1n, 0, FG);
      x2n = __shfl_sync(0xffffffffu, x2n, 0, FG);
      unew = __shfl_sync(0xffffffffu, unew, 0, FG);
      }
      // ================= lift(x+) =================
      double yl;
      if (MLP) {
        // x+ -> the unit's layer-0 block (k-major); the barrier also says that both warps of the unit
        // are done with the scratch their activations are about to overwrite
        if (l < 4) in0[act_index(l, sc & 7)] = (l == 0) ? x1n : ((l == 1) ? x2n : 0.0);
        quarter_barrier(qbar);
        if (TIMED) tq += clock64() - c0, c0 = clock64();
        if (!(KMPC_DBG_SKIP(a) & 2))
          lift_unit<2>(a.p, in0, region, region + a.sm.actbuf, yout, FNZ, wsm, wl ^ ((quarter >> 1) & 1), tid & 31, qbar);
        yl = yout[(sc & 7) * FNZ + l];
        if (c.lift_mode != KMPC_LIFT_RAW) yl -= a.p.z0[l];
        if (TIMED) tl += clock64() - c0, c0 = clock64();
      } else {
        if (TIMED) tq += clock64() - c0, c0 = clock64();
        const double xx[2] = {x1n, x2n};
        yl = rbf_thinplate(xx, b.cx + l * 2, 2, c.lift_mode);
        if (TIMED) tl += clock64() - c0, c0 = clock64();
      }
      // ================= RLS (duffing.py:927-953, 965-984) =================
      if (UPDATE && !(KMPC_DBG_SKIP(a) & 4)) {
        double zv[8];
        group_gather(ex, l, zl, zv);
        const double u = unew;
        // w = P v (rows l and 8), rrow = v'P
        double w = 0.0;
#pragma unroll
        for (int j = 0; j < FNZ; ++j) w = fma(Pr[j], zv[j], w);
        w = fma(Pr[8], u, w);
        double p8v[8];
        group_gather(ex + 8, l, P8l, p8v);
        double w8 = 0.0;
#pragma unroll
        for (int j = 0; j < FNZ; ++j) w8 = fma(p8v[j], zv[j], w8);
        w8 = fma(P88, u, w8);
        double ch[8];
#pragma unroll
        for (int j = 0; j < FNZ; ++j) ch[j] = zl * Pr[j];
        double rr = chunk_reduce(red, l, ch);
        rr = fma(u, P8l, rr);                    // rrow[l]
        double c8[8];
        group_gather(ex + 16, l, zl * Pr[8], c8);
        double rr8 = c8[0];
#pragma unroll
        for (int j = 1; j < FNZ; ++j) rr8 += c8[j];
        rr8 = fma(u, P88, rr8);                  // rrow[8]
        double rv[8];
        group_gather(ex + 24, l, rr, rv);
        double vPv = 0.0;                        // (v'P) v, duffing.py:934
#pragma unroll
        for (int j = 0; j < FNZ; ++j) vPv = fma(rv[j], zv[j], vPv);
        vPv = fma(rr8, u, vPv);
        const double denom = lam + vPv;
        if (lam == 1.0) {  // x / 1 == x exactly; one reciprocal instead of eleven divisions
          const double wd = w / denom, w8d = w8 / denom;
#pragma unroll
          for (int j = 0; j < FNZ; ++j) Pr[j] = fma(-wd, rv[j], Pr[j]);
          Pr[8] = fma(-wd, rr8, Pr[8]);
          P8l = fma(-w8d, rr, P8l);
          P88 = fma(-w8d, rr8, P88);
        } else {           // Koopman_update.m:270 forgetting factor
#pragma unroll
          for (int j = 0; j < FNZ; ++j) Pr[j] = Pr[j] / lam - (w * rv[j]) / lam / denom;
          Pr[8] = Pr[8] / lam - (w * rr8) / lam / denom;
          P8l = P8l / lam - (w8 * rr) / lam / denom;
          P88 = P88 / lam - (w8 * rr8) / lam / denom;
        }
#pragma unroll
        for (int j = 0; j < FNZ; ++j) KAr[j] = fma(yl, zv[j], KAr[j]);
        KAr[8] = fma(yl, u, KAr[8]);
        // [A B] = K_A P: all-gather the new P through the scratch (81 doubles)
        double* Pm = scr;
#pragma unroll
        for (int j = 0; j < FNV; ++j) Pm[l * 9 + j] = Pr[j];
        Pm[72 + l] = P8l;
        if (l == 0) Pm[80] = P88;
        __syncwarp();
#pragma unroll 1
        for (int j = 0; j < FNV; j += 3) {   // columns j .. j + 2 of [A B] (three chains in flight); column 8 is B (HF[64 + l])
          double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int k = 0; k < FNV; ++k) {
            s0 = fma(KAr[k], Pm[k * 9 + j], s0);
            s1 = fma(KAr[k], Pm[k * 9 + j + 1], s1);
            s2 = fma(KAr[k], Pm[k * 9 + j + 2], s2);
          }
          HF[j * 8 + l] = s0;
          HF[(j + 1) * 8 + l] = s1;
          HF[(j + 2) * 8 + l] = s2;
        }
        __syncwarp();
        if (upc) {
          double wq = 0.0;
#pragma unroll
          for (int j = 0; j < FNZ; ++j) wq = fma(Qr[j], zv[j], wq);
#pragma unroll
          for (int j = 0; j < FNZ; ++j) ch[j] = zl * Qr[j];
          const double rq = chunk_reduce(red, l, ch);
          double rqv[8];
          group_gather(ex, l, rq, rqv);
          double zQz = 0.0;
#pragma unroll
          for (int j = 0; j < FNZ; ++j) zQz = fma(rqv[j], zv[j], zQz);
          const double dq = 1.0 + zQz;
          const double wqd = wq / dq;
#pragma unroll
          for (int j = 0; j < FNZ; ++j) Qr[j] = fma(-wqd, rqv[j], Qr[j]);
          const bool skipx = (t == 0 && a.first && c.skip_first_barx);  // Tank_System.m:252-254
          if (!skipx) {
            const double xc0 = c.c_pairs_next ? x1n : x1, xc1 = c.c_pairs_next ? x2n : x2;
            Xc0 = fma(xc0, zl, Xc0);
            Xc1 = fma(xc1, zl, Xc1);
          }
          // C = bar_X bar_Q: the QP reads it every step only with y = C z; with y = z it is an output of
          // the launch and is formed once, from the final bar_X and bar_Q (same arithmetic)
          if (OUT == KMPC_OUT_C || t == a.T - 1) {
#pragma unroll
            for (int j = 0; j < FNZ; ++j) ch[j] = Xc0 * Qr[j];
            Cc0 = chunk_reduce(red, l, ch);
#pragma unroll
            for (int j = 0; j < FNZ; ++j) ch[j] = Xc1 * Qr[j];
            Cc1 = chunk_reduce(red, l, ch);
          }
        }
      }
      zl = yl;
      x1 = x1n;
      x2 = x2n;
      uprev = unew;
      if (TIMED) tr += clock64() - c0;
    }

    // ---- registers -> state ----
    if (valid) {
      b.z[s * 8 + l] = zl;
      if (l == 0) {
        b.x[s * 2] = x1;
        b.x[s * 2 + 1] = x2;
        b.u_prev[s] = uprev;
        if (b.status && status) b.status[s] |= status;
        if (a.d.wset) reinterpret_cast<uint2*>(a.d.wset)[s] = make_uint2(wlo, whi);
      }
      if (UPDATE && a.T > 0) {
#pragma unroll
        for (int j = 0; j < FNZ; ++j) b.A[s * 64 + l * 8 + j] = HF[j * 8 + l];
        b.B[s * 8 + l] = HF[64 + l];
#pragma unroll
        for (int j = 0; j < FNV; ++j) {
          b.KA[s * 72 + l * 9 + j] = KAr[j];
          b.P[s * 81 + l * 9 + j] = Pr[j];
        }
        b.P[s * 81 + 72 + l] = P8l;
        if (l == 0) b.P[s * 81 + 80] = P88;
        if (upc) {
          b.C[s * 16 + l] = Cc0;
          b.C[s * 16 + 8 + l] = Cc1;
#pragma unroll
          for (int j = 0; j < FNZ; ++j) b.barQ[s * 64 + l * 8 + j] = Qr[j];
          b.barX[s * 16 + l] = Xc0;
          b.barX[s * 16 + 8 + l] = Xc1;
        }
      }
    }
  }
  if (TIMED && tid == 0) {
    a.timing[blockIdx.x * 4 + 0] = tq;
    a.timing[blockIdx.x * 4 + 1] = tl;
    a.timing[blockIdx.x * 4 + 2] = tr;
    a.timing[blockIdx.x * 4 + 3] = clock64() - t_begin;
  }
}

typedef void (*FusedKernel)(const FusedArgs);

template <bool TIMED>
static FusedKernel pick_fused(int out_mode, bool update, bool mlp) {
  if (out_mode == KMPC_OUT_IDENTITY) {
    if (update) return mlp ? fused_loop_kernel<KMPC_OUT_IDENTITY, true, true, TIMED>
                           : fused_loop_kernel<KMPC_OUT_IDENTITY, true, false, TIMED>;
    return mlp ? fused_loop_kernel<KMPC_OUT_IDENTITY, false, true, TIMED>
               : fused_loop_kernel<KMPC_OUT_IDENTITY, false, false, TIMED>;
  }
  if (update) return mlp ? fused_loop_kernel<KMPC_OUT_C, true, true, TIMED>
                         : fused_loop_kernel<KMPC_OUT_C, true, false, TIMED>;
  return mlp ? fused_loop_kernel<KMPC_OUT_C, false, true, TIMED> : fused_loop_kernel<KMPC_OUT_C, false, false, TIMED>;
}

// Can the fused kernel run this loop?  (cfg.path = KMPC_PATH_GENERIC forces the generic three-kernel
// path: an explicit, per-context choice -- no environment variable is read on the launch path.)
bool fused_eligible(const kmpc_loop_config& c, const kmpc_encoder* enc) {
  if (c.path == KMPC_PATH_GENERIC) return false;
  if (c.nz != FNZ || c.N != FN || c.n != 2 || c.du_aug) return false;
  if (c.out_mode != KMPC_OUT_IDENTITY && c.out_mode != KMPC_OUT_C) return false;
  if (c.update && !(c.rls_flags & KMPC_RLS_UPDATE_C) && c.out_mode == KMPC_OUT_C) {
    // C is then a frozen input: fine, it is simply never rewritten
  }
  if (c.lift_kind == KMPC_LIFTKIND_MLP) {
    if (!enc || enc->smem_bytes <= 0) return false;
    if (c.lift_mode == KMPC_LIFT_STACK) return false;
    if (enc->p.dims[0] != 2 || enc->p.dims[enc->p.n_layers] != FNZ) return false;
    if (enc->p.actw * kActPitch < 128) return false;            // split-K partials live in a ping-pong buffer
    const FusedSmem L = fused_smem_layout(&enc->p);
    if (L.total_bytes > enc->max_smem_optin) return false;
  } else if (c.lift_kind != KMPC_LIFTKIND_RBF) {
    return false;
  }
  return true;
}

// Launch T steps.  `timing` (nullable, device, [grid][4] long long) selects the TIMED build.
int fused_launch(const LoopDev& d, const kmpc_encoder* enc, int64_t step0, int T, int first,
                 long long* timing, int* grid_out, cudaStream_t st) {
  const kmpc_loop_config& c = d.c;
  const bool mlp = c.lift_kind == KMPC_LIFTKIND_MLP;
  FusedArgs a;
  a.d = d;
  if (mlp) a.p = enc->p;
  else memset(&a.p, 0, sizeof(a.p));
  a.sm = fused_smem_layout(mlp ? &enc->p : nullptr);
  a.step0 = step0;
  a.T = T;
  a.first = first;
  a.num_tiles = (c.S + kTileS - 1) / kTileS;
  a.timing = timing;
#ifdef KMPC_PROFILING
  {
    const char* e = getenv("KMPC_FUSED_SKIP");
    a.dbg_skip = e ? atoi(e) : 0;
  }
#endif
  int dev = 0, sms = 148;
  KMPC_CUDA(cudaGetDevice(&dev));
  KMPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t cap = mlp ? sms : (int64_t)sms * 4;
  const unsigned grid = (unsigned)(a.num_tiles < cap ? a.num_tiles : cap);
  FusedKernel k = timing ? pick_fused<true>(c.out_mode, c.update != 0, mlp)
                         : pick_fused<false>(c.out_mode, c.update != 0, mlp);
  KMPC_CUDA(ensure_smem(k, a.sm.total_bytes));
  k<<<grid, kMmaThreads, a.sm.total_bytes, st>>>(a);
  KMPC_AFTER_LAUNCH();
  if (grid_out) *grid_out = (int)grid;
  return KMPC_OK;
}

}  // namespace kmpc
