/* kmpc.h -- C ABI of libkmpc.so: B200 (sm_100a) kernels for the closed-loop Koopman-MPC hot path.
 *
 * The reference (MichaelMillerCSU/Koopman-online-updated-MPC) has no FFI or plugin interface: its
 * "API" is the expressions at the call sites of duffing.py / vanderpol.py / duffing_RBF.py /
 * Tank_System.m.  Each entry point below replaces one of those expressions, batched over a leading
 * scenario axis S.  INTEGRATION.md shows the ctypes stubs a maintainer of the reference would add.
 *
 * Conventions
 *   - every `double*` / `int*` marked [dev] is a DEVICE pointer owned by the caller (e.g.
 *     torch.Tensor.data_ptr()); float64, C-contiguous, scenario-major (S, rows, cols);
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *     every call is asynchronous on it and performs no host<->device copies, except the
 *     *_create functions which upload host weights once;
 *   - return value: 0 = OK, negative = argument / CUDA error (kmpc_strerror); numerical trouble
 *     is reported per scenario in `status` arrays (KMPC_STATUS_* bits), never by aborting a batch;
 *   - no CPU fallback exists: without a CUDA device every compute call returns KMPC_ERR_CUDA.
 *   - input dimension m = 1 (all reference systems are single-input).
 */
#ifndef KMPC_H
#define KMPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMPC_VERSION 100

/* error codes */
#define KMPC_OK 0
#define KMPC_ERR_ARG (-1)       /* bad argument (null pointer, dimension out of range) */
#define KMPC_ERR_CUDA (-2)      /* CUDA runtime / launch error (see kmpc_last_cuda_error) */
#define KMPC_ERR_UNSUPPORTED (-3)
#define KMPC_ERR_ALLOC (-4)

/* per-scenario status bits */
#define KMPC_STATUS_MAXITER 1   /* active-set iteration cap hit */
#define KMPC_STATUS_NONFINITE 2 /* non-finite value in the result */
#define KMPC_STATUS_PIVOT 4     /* non-positive Cholesky pivot (matrix not SPD to working precision) */

/* limits of the generic kernels */
#define KMPC_MAX_NZ 16          /* lifted dimension seen by RLS / QP (incl. du-augmentation) */
#define KMPC_MAX_HORIZON 64
#define KMPC_MAX_LAYERS 8
#define KMPC_MAX_WIDTH 128      /* widest MLP layer */

const char* kmpc_strerror(int code);
const char* kmpc_last_cuda_error(void);
int kmpc_version(void);
/* number of kernel launches issued by this library in this process (bench.py's gpu_launches) */
int64_t kmpc_launch_count(void);

/* ------------------------------------------------------------------ stage 1: lifting ---------
 * theta_E encoder: duffing.py:21-29 (`net.Encoder`, call sites l.153,155,764,847,884),
 * vanderpol.py:154,672,679,760,769,805; Encoder_Duffing.m:3-6, Encoder_VDP.m:3-6, Encoder_Tank.m:3-5.
 * W[l] is HOST memory, nn.Linear layout (dims[l+1], dims[l]) row-major; b[l] HOST (dims[l+1]).
 * Input width dims[0] <= 16: the Decoder half (duffing.py:30-38, 8 -> 100 -> 100 -> 100 -> 2) is the
 * same kind of handle (used by the training-loss evaluation); nets with dims[0] > 4 run on the
 * CTA-wide fp64 tensor kernel and have no KMPC_PREC_TC path.                                      */
typedef struct kmpc_encoder kmpc_encoder;

#define KMPC_LIFT_RAW 0    /* theta(x)                        duffing.py:764                      */
#define KMPC_LIFT_OFFSET 1 /* theta(x) - theta(0)             Koopman_update_Tracking_Lift.m:65   */
#define KMPC_LIFT_STACK 2  /* [x; theta(x)] - [0; theta(0)]   Koopman_update.m:67                 */

int kmpc_encoder_create(kmpc_encoder** out, const double* const* W, const double* const* b,
                        const int* dims, int n_layers, void* stream);
int kmpc_encoder_destroy(kmpc_encoder* enc);
/* output width for a lift mode (nz, or n + nz for STACK) */
int kmpc_encoder_out_dim(const kmpc_encoder* enc, int lift_mode);
/* x [dev] (S, n) -> z [dev] (S, out_dim) */
int kmpc_encode(const kmpc_encoder* enc, const double* x, double* z, int64_t S, int lift_mode,
                void* stream);

/* Precision of the EDMD-side lift (duffing.py:152-164, the 20 000 encoder calls ahead of the regression):
 *   KMPC_PREC_FP64  fp64 tensor path (mma.sync m8n8k4 f64), bit-compatible with the closed loop's lift;
 *   KMPC_PREC_TC    Blackwell tensor path: tcgen05.mma kind::f16 on bf16 x 3 split operands (six piece
 *                   products), fp32 accumulation in TMEM, weights staged by TMA tensor copies; lifted
 *                   states agree with fp64 to ~1e-7 relative (tests/test_tc_lift.py), the Gram pack is
 *                   still accumulated in fp64.  Nets with hidden width <= 112, out <= 16, in <= 4;
 *                   otherwise KMPC_ERR_UNSUPPORTED (kmpc_encoder_has_tc tells in advance).           */
#define KMPC_PREC_FP64 0
#define KMPC_PREC_TC 1
int kmpc_encoder_has_tc(const kmpc_encoder* enc);
int kmpc_encode_ex(const kmpc_encoder* enc, const double* x, double* z, int64_t S, int lift_mode,
                   int precision, void* stream);

/* thin-plate RBF lift: duffing_RBF.py:20-23 (variant 0: d^2 log(d + 1e-4)); rbf.m:24-29
 * (variant 1: r2 log sqrt(r2), 0 at r = 0).  x [dev] (S, n), cx [dev] (nz, n) -> z [dev] (S, nz) */
#define KMPC_RBF_PYTHON 0
#define KMPC_RBF_MATLAB 1
int kmpc_rbf_lift(const double* x, const double* cx, double* z, int64_t S, int n, int nz,
                  int variant, void* stream);

/* ------------------------------------------------------------------ stage 2: EDMD ------------
 * duffing.py:167-177 ([A B] = PHIY pinv([PHIX;U]), C = X pinv(PHIX)); Gram form Tank_System.m:93-100.
 * pack [dev] layout (doubles): G = V V' (nv*nv) | Aq = PHIY V' (nz*nv) | XV = X V' (n*nv) | count,
 * V = [PHIX; U], nv = nz + 1.  kmpc_gram_accumulate ADDS into pack (zero it first); snapshots are
 * row-major (M, nz), (M, nz), (M, 1), (M, n).  The pack is what is all-reduced across GPUs.      */
int64_t kmpc_gram_pack_len(int nz, int n);
int kmpc_gram_accumulate(const double* psi, const double* psi_next, const double* u,
                         const double* x, int64_t M, int nz, int n, double* pack, void* stream);
/* fused lift + Gram: reads raw snapshots x, y (M, n), u (M, 1), never materialises PHIX/PHIY */
int kmpc_gram_from_snapshots(const kmpc_encoder* enc, int lift_mode, const double* x,
                             const double* y, const double* u, int64_t M, double* pack,
                             void* stream);
/* trajectory-aware variant for CONSECUTIVE trajectory-major snapshots (data_generate.py:63-74 /
 * kmpc_generate_snapshots: y of snapshot j is x of snapshot j + 1 of the same trajectory): one
 * encode per state (n_step + 1 per trajectory) instead of two per snapshot; same pack.          */
int kmpc_gram_from_trajectories(const kmpc_encoder* enc, int lift_mode, const double* x,
                                const double* y, const double* u, int64_t n_traj, int n_step,
                                double* pack, void* stream);
/* the same two with an explicit lift precision (KMPC_PREC_*).  One lift workspace lives in the encoder
 * handle: concurrent calls on the SAME handle (threads or streams) are serialised by the library.   */
int kmpc_gram_from_snapshots_ex(const kmpc_encoder* enc, int lift_mode, int precision, const double* x,
                                const double* y, const double* u, int64_t M, double* pack,
                                void* stream);
int kmpc_gram_from_trajectories_ex(const kmpc_encoder* enc, int lift_mode, int precision,
                                   const double* x, const double* y, const double* u, int64_t n_traj,
                                   int n_step, double* pack, void* stream);
#define KMPC_C_PYTHON 0 /* C = (X PHIX')(PHIX PHIX')^-1            duffing.py:177        */
#define KMPC_C_JOINT 1  /* C = block of [PHIY;X] V' (V V')^-1      Tank_System.m:96-100  */
/* A [dev] (nz,nz), B [dev] (nz,1), C [dev] (n,nz), status [dev] (1) */
int kmpc_edmd_solve(const double* pack, int nz, int n, int c_variant, double* A, double* B,
                    double* C, int* status, void* stream);

/* ------------------------------------------------------------------ stage 3: online update ---
 * duffing.py:900,927-953,965-984; vanderpol.py:872-895; Koopman_update.m:258-278 (lambda);
 * Tank_System.m:234-263.  State [dev], updated in place: KA (S,nz,nv), P (S,nv,nv),
 * barX (S,n,nz), barQ (S,nz,nz).  Sample: z (S,nz), u (S,1), y (S,nz), xc (S,n) = the state paired
 * with z in the C regression.  Outputs A (S,nz,nz), B (S,nz,1), C (S,n,nz).                      */
#define KMPC_RLS_UPDATE_C 1      /* update bar_X/bar_Q and emit C (off: Koopman_update.m)         */
#define KMPC_RLS_SKIP_BARX 2     /* do not accumulate bar_X this call (Tank_System.m:252-254)     */
int kmpc_rls_update(double* KA, double* P, double* barX, double* barQ, const double* z,
                    const double* u, const double* y, const double* xc, double* A, double* B,
                    double* C, int64_t S, int nz, int n, double lambda, int flags, void* stream);

/* ------------------------------------------------------------------ stage 4: MPC QP ----------
 * Replaces `optimize.minimize(costFunction, zeros(N), bounds=...).x` (duffing.py:540-581,776-778)
 * and `quadprog(2H, f, ..., lb, ub)` (Tank_System.m:128-159,188): condensed build
 * H = G'QG + R, f = 2 G'Q(F z0 - r), then an exact primal active-set solve of
 *   min U'HU + f'U  s.t. lb <= U <= ub.
 * A (S|1,nz,nz), B (S|1,nz), Cy (S|1,ny,nz) (model shared by all scenarios when
 * KMPC_QP_SHARED_MODEL), z0 (S,nz), r (S,ny) constant over the horizon or (S,N,ny) with
 * KMPC_QP_R_FULL, lb/ub (S,N), PN optional terminal weight (S|1,ny,ny) replacing the last q*I.
 * Outputs u0 (S), Ufull (S,N) nullable, status (S) nullable.                                     */
#define KMPC_QP_SHARED_MODEL 1
#define KMPC_QP_R_FULL 2
#define KMPC_QP_CY_IDENTITY 4 /* Cy = I (ny == nz), Cy pointer ignored  vanderpol.py:456-459 */
int kmpc_qp_first_move(const double* A, const double* B, const double* Cy, const double* z0,
                       const double* r, const double* lb, const double* ub, const double* PN,
                       double q, double rw, int N, int ny, int nz, int64_t S, int flags,
                       double* u0, double* Ufull, int* status, int max_iter, double tol,
                       void* stream);

/* ------------------------------------------------------------------ plant --------------------
 * duffing.py:250-261 (RK4, h = 0.05), Koopman_update.m:21-25 (k4 uses k1), Tank_System.m:9-10,211.
 * params (S,5): POLY2  x1' = p0 x2 ; x2' = p1 x2 + p2 x1 + p3 x1^3 + p4 x1^2 x2 + u
 *               TANK   x1+ = x1 - p0 sqrt(x1) + p1 u ; x2+ = x2 + p2 sqrt(x1) - p3 sqrt(x2) ; >= 0 */
#define KMPC_PLANT_POLY2 0
#define KMPC_PLANT_TANK 1
#define KMPC_RK4_PYTHON 0
#define KMPC_RK4_MATLAB 1
int kmpc_plant_step(const double* x, const double* u, const double* params, double* xnext,
                    int64_t S, int kind, int rk4_variant, double h, void* stream);

/* ------------------------------------------------------------------ fused closed loop --------
 * One scenario-step = duffing.py:823-992 loop body (lift -> QP -> plant -> lift -> RLS), or the
 * Tank_System.m:170-291 body with du_aug.  All state lives in caller-owned device buffers.      */
#define KMPC_OUT_C 0        /* y = C z, ny = n                                  */
#define KMPC_OUT_IDENTITY 1 /* y = z,   ny = nz          vanderpol.py:456       */
#define KMPC_OUT_C_ROW 2    /* y = (C z)[out_row], ny=1  Tank_System.m:113      */
#define KMPC_LIFTKIND_MLP 0
#define KMPC_LIFTKIND_RBF 1
#define KMPC_PATH_AUTO 0
#define KMPC_PATH_GENERIC 1

typedef struct kmpc_loop_config {
  int64_t S;
  int nz;            /* lift dimension (RLS dimension); QP sees nz + du_aug */
  int n;             /* plant state dimension (2) */
  int N;             /* horizon */
  int out_mode;      /* KMPC_OUT_* */
  int out_row;
  int du_aug;        /* velocity form, Tank_System.m:110-113 */
  int update;        /* 0 = frozen model (duffing.py:738-805), 1 = online RLS (l.823-992) */
  int rls_flags;     /* KMPC_RLS_UPDATE_C */
  int c_pairs_next;  /* 1: bar_X += x+ z' (python); 0: bar_X += x z' (tank) */
  int skip_first_barx;
  int shared_model;  /* frozen model shared by all scenarios: A,B,C are (1,...) */
  int lift_kind;     /* KMPC_LIFTKIND_* */
  int lift_mode;     /* KMPC_LIFT_* (MLP) or KMPC_RBF_* (RBF) */
  int plant_kind;
  int rk4_variant;
  int first_post_step; /* first step index integrated with params_post (python 102, matlab 100) */
  int max_iter;
  double h;
  double q, rw;
  double lb, ub;       /* move bounds */
  double u_lb, u_ub;   /* absolute input bounds on the first move when du_aug */
  double lambda;
  double p0, q0;       /* RLS restart: P = p0 I, bar_Q = q0 I at the first update (step_index == 0
                          and rls_started == 0); ignored when the caller warm-starts the state */
  double tol;
  int path;            /* KMPC_PATH_AUTO: persistent fused kernel when the shape allows it;
                          KMPC_PATH_GENERIC: always the per-step qp_plant -> lift -> rls kernels (cross-check) */
  int qp_cold;         /* generic kernels: 0 = warm start from the previous step's moves, primal-dual sweeps then
                          the primal method (same minimiser as a cold start); 1 = cold-start every QP like the
                          reference (duffing.py:634: pastRes is never written back); 2 = warm start, primal
                          method only (round-1 behaviour, kept for A/B timing); 3 = like 0 with DAMPED sweeps on the
                          warp-per-scenario shapes (every violated bound is clipped but only the most negative
                          multiplier is released per sweep, up to 40 sweeps): same minimiser, fewer factorisations
                          where the plain sweeps cycle (Tank: -18 %); where the Hessian is numerically singular
                          (cond > 1e16, flagged KMPC_STATUS_PIVOT) the iterate it stops at may differ from mode 0's */
} kmpc_loop_config;

typedef struct kmpc_loop_buffers {     /* all [dev] */
  double* x;          /* (S,n)   current plant state, updated in place */
  double* z;          /* (S,nz)  lift(x), kept consistent by the library */
  double* u_prev;     /* (S)     last applied input */
  double* A;          /* (S|1,nz,nz) */
  double* B;          /* (S|1,nz)    */
  double* C;          /* (S|1,n,nz)  */
  double* KA;         /* (S,nz,nz+1) RLS state (update only) */
  double* P;          /* (S,nz+1,nz+1) */
  double* barX;       /* (S,n,nz) */
  double* barQ;       /* (S,nz,nz) */
  const double* r;    /* (S,ny) reference, constant over the horizon */
  const double* params_pre;   /* (S,5) */
  const double* params_post;  /* (S,5) */
  const double* cx;   /* (nz,n) RBF centres (lift_kind RBF) */
  double* log_x;      /* nullable (T_cap,S,n)  x after each step */
  double* log_u;      /* nullable (T_cap,S)    */
  int* status;        /* nullable (S) OR-accumulated KMPC_STATUS_* */
  int64_t log_capacity; /* T_cap */
} kmpc_loop_buffers;

/* Streams: kmpc_ctx_create / kmpc_ctx_reset queue small memsets on the stream they are given and
 * record an event; kmpc_closed_loop_steps* waits on that event, so create / reset / steps may use
 * different streams.  The caller's own writes to the buffers must be ordered before the steps call
 * by the caller.  A ctx may be used from one host thread at a time.                              */
typedef struct kmpc_ctx kmpc_ctx;
int kmpc_ctx_create(kmpc_ctx** out, const kmpc_loop_config* cfg, const kmpc_loop_buffers* buf,
                    const kmpc_encoder* enc /* nullable for RBF */, int rls_started,
                    void* stream);
int kmpc_ctx_destroy(kmpc_ctx* ctx);
/* run T scenario-steps for all S scenarios; continues from the ctx's step index */
int kmpc_closed_loop_steps(kmpc_ctx* ctx, int T, void* stream);
int64_t kmpc_ctx_step_index(const kmpc_ctx* ctx);
/* start a new episode on the same buffers: step index <- 0, RLS restart pending again unless
 * rls_started, QP warm-start memory cleared.  The caller rewrites x, z, u_prev, A, B, C (and the
 * RLS state when warm-starting) -- the loop `for i in range(maxStep)` begins again (duffing.py:823) */
int kmpc_ctx_reset(kmpc_ctx* ctx, int rls_started, void* stream);
/* 1 when kmpc_closed_loop_steps runs as ONE persistent fused kernel per call (nz = 8, N = 10 loops),
 * 0 when it runs the generic qp_plant -> lift -> rls kernels once per step */
int kmpc_ctx_is_fused(const kmpc_ctx* ctx);
/* same as kmpc_closed_loop_steps but also returns, after synchronising, the device time in
 * milliseconds spent in ms[0] = QP+plant, ms[1] = lift, ms[2] = RLS (bench.py's roofline):
 * CUDA events around every kernel on the generic path (T <= 1024), per-phase clock64() sums
 * scaled to the event-timed launch on the fused path */
int kmpc_closed_loop_steps_timed(kmpc_ctx* ctx, int T, void* stream, float* ms);

/* ------------------------------------------------------------------ snapshot generator -------
 * data_generate.py:17-74 (duffing_generate) / 82-152 (vanderpol_generate): n_traj trajectories,
 * n_step vectorised RK4 steps (h = 0.05, zero-order-hold input), outputs re-ordered trajectory-major
 * (l.63-74) -- snapshot-major here: X, Y (M, n = 2), U (M), M = n_traj * n_step, snapshot
 * traj * n_step + j = step j of trajectory traj.  The random draws stay with the caller
 * (reference: u0 = 4 rand(N, N_Traj) - 2 then x0 = 4 rand(n, N_Traj) - 2 from numpy's global
 * stream): x0 [dev] (n_traj, 2), u0 [dev] (n_step, n_traj) in the reference's own layout;
 * params [dev] (5) as for kmpc_plant_step.                                                        */
int kmpc_generate_snapshots(const double* x0, const double* u0, const double* params, int plant_kind,
                            int rk4_variant, double h, int64_t n_traj, int n_step, double* X, double* Y,
                            double* U, void* stream);

/* ------------------------------------------------------------------ open-loop predictor ------
 * duffing.py:290-343 / vanderpol.py:292-348: along T consecutive snapshots of each of n_seq
 * sequences (sequence s starts at snapshot s * seq_stride >= 1, windows may overlap; the reference
 * checks one: n_seq = 1, T = plotTime) the lifted state restarts from the TRUE lifted state psi every reset_every (10)
 * steps and follows z+ = A z + B u in between; logged BEFORE the propagation: decoder_X (n_seq,T,nz)
 * = z and test_Y (n_seq,T,n) = C z.  psi [dev] (M, nz) = lift of the snapshots (kmpc_encode /
 * kmpc_rbf_lift), x [dev] (M, n), u [dev] (M).  rmse (nullable, [dev] (n_seq)):
 * || (test_Y[:, rmse_row] - x[:T, rmse_row]) / T ||_2 (duffing.py:341 row 0, vanderpol.py:346 row 1). */
int kmpc_open_loop_predict(const double* psi, const double* x, const double* u, const double* A,
                           const double* B, const double* C, int nz, int n, int64_t n_seq, int T,
                           int64_t seq_stride, int reset_every, int rmse_row, double* decoder_X,
                           double* test_Y, double* rmse, void* stream);

/* ------------------------------------------------------------------ training-loss windows ----
 * duffing.py:179-235 (and the loop body of DeepLearning_KoopmanControl_Approach3.py:462-563): for window
 * w starting at snapshot k = k0 + w * stride, with zpred (W,T,nz) the linear rollout of the window
 * (kmpc_open_loop_predict with seq_stride = stride, reset_every = T: zpred[w][0] = psi_k) and
 * xdec (W,T,n) = Decoder(zpred) (a second kmpc_encoder holding the decoder half):
 *   out[w] = { ||xdec[w][0] - x_k||^2,  sum_{p=1..T-1} ||zpred[w][p] - psi_{k+p}||^2,
 *              sum_{p=1..T-1} ||x_{k+p} - xdec[w][p]||^2 }        (criterion = MSELoss(reduction='sum'))
 * The reference's own accumulation over windows (l.179, 221-233) is host arithmetic on these sums.   */
int kmpc_window_losses(const double* psi, const double* x, const double* zpred, const double* xdec, int nz,
                       int n, int64_t W, int T, int64_t k0, int64_t stride, double* out, void* stream);

/* ------------------------------------------------------------------ roofline denominators ----
 * MEASURED_PEAKS.json carries no fp64 number: measure this GPU's fp64 tensor-path
 * (mma.sync.m8n8k4.f64) and CUDA-core (DFMA) peaks, TFLOP/s, ~25 ms each; synchronises. */
int kmpc_measure_fp64_peak(double* dmma_tflops, double* dfma_tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KMPC_H */
