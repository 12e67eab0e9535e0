// closed_loop.cu -- the fused closed-loop scenario step (the path the headline metric times).
//
// One scenario-step (duffing.py:823-992; Tank_System.m:170-291 with du_aug):
//     z = lift(x) -> condensed box-QP -> u -> x+ = plant(x, u) -> y = lift(x+) -> RLS(z,u,y) -> A,B,C
// realised per step as  [qp_plant kernel] -> [lift kernel] -> [rls kernel]  on one stream, with all
// state resident in caller-owned device buffers (so a step's HBM traffic is exactly the
// algorithmic bytes of SURVEY.md 8d: RLS state read+write, A/B/C write, x/u/z).
#include <new>
#include <vector>

#include "common.cuh"
#include "loopbody.cuh"

namespace kmpc {

// fused.cu: the persistent kernel for nz = 8, N = 10 loops
bool fused_eligible(const kmpc_loop_config& c, const kmpc_encoder* enc);
int fused_launch(const LoopDev& d, const kmpc_encoder* enc, int64_t step0, int T, int first,
                 long long* timing, int* grid_out, cudaStream_t st);

constexpr int kLoopThreads = 128;      // default block
constexpr int kLoopThreadsMax = 256;   // large QP workspaces: one bigger block per SM (see select_launch)

// G lanes per scenario; NZ/N/OUT/DU > 0 give a compile-time QP shape (loops unrolled, indices
// folded), NZ == 0 reads the shape from the config at run time.
template <int G, int NZ, int N, int OUT, int DU>
__global__ void __launch_bounds__(kLoopThreadsMax)
loop_qp_plant_kernel(LoopDev d, int64_t step, int64_t log_slot) {
  extern __shared__ double smem[];
  pdl_wait();  // everything this kernel reads (A, B, C, z, x) comes from the previous kernels
  pdl_launch_dependents();
  const int group = threadIdx.x / G;
  int64_t s = (int64_t)blockIdx.x * (blockDim.x / G) + group;
  const bool valid = s < d.c.S;
  if (!valid) s = d.c.S - 1;
  LoopShape sh;
  if (NZ > 0) {
    sh.nz = NZ; sh.N = N; sh.out_mode = OUT; sh.du_aug = DU;
  } else {
    sh = loop_shape(d.c);
  }
  const int nzq = sh.nz + sh.du_aug;
  const bool identity = sh.out_mode == KMPC_OUT_IDENTITY;
  const int ny = identity ? nzq : (sh.out_mode == KMPC_OUT_C ? 2 : 1);
  loop_qp_plant_scenario<G, (N > 0 ? N : KMPC_MAX_HORIZON), (NZ > 0 ? NZ + DU : 0)>(d, sh, s, valid, step, log_slot,
                            smem + (size_t)group * qp_ws_doubles(nzq, ny, sh.N, identity));
}

template <int G, int NZ>
__global__ void __launch_bounds__(kLoopThreads)
loop_rls_kernel(LoopDev d, int first) {
  extern __shared__ double smem[];
  const int group = threadIdx.x / G;
  int64_t s = (int64_t)blockIdx.x * (blockDim.x / G) + group;
  const bool valid = s < d.c.S;
  if (!valid) s = d.c.S - 1;
  const int nz = NZ > 0 ? NZ : d.c.nz;
  loop_rls_scenario<G>(d, nz, s, valid, first, smem + (size_t)group * rls_ws_doubles(nz, 2));
}

typedef void (*QpKernel)(LoopDev, int64_t, int64_t);
typedef void (*RlsKernel)(LoopDev, int);
struct LoopLaunch {
  QpKernel qp;
  RlsKernel rls;
  int qp_g, rls_g;          // lanes per scenario
  int qp_spb, rls_spb;      // scenarios per block
  int qp_smem, rls_smem;    // dynamic shared memory per block (bytes)
};

static bool select_launch(const kmpc_loop_config& c, LoopLaunch* L) {
  const int nzq = c.nz + (c.du_aug ? 1 : 0);
  const bool identity = c.out_mode == KMPC_OUT_IDENTITY;
  const int ny = identity ? nzq : (c.out_mode == KMPC_OUT_C ? 2 : 1);
  L->qp_g = 32;
  if (c.nz == 8 && c.N == 10 && !c.du_aug && c.out_mode == KMPC_OUT_IDENTITY) {
    L->qp = loop_qp_plant_kernel<16, 8, 10, KMPC_OUT_IDENTITY, 0>;   // vanderpol.py
    L->qp_g = 16;
  } else if (c.nz == 8 && c.N == 10 && !c.du_aug && c.out_mode == KMPC_OUT_C) {
    L->qp = loop_qp_plant_kernel<16, 8, 10, KMPC_OUT_C, 0>;          // duffing.py, duffing_RBF.py
    L->qp_g = 16;
  } else if (c.nz == 10 && c.N == 20 && c.du_aug && c.out_mode == KMPC_OUT_C_ROW) {
    L->qp = loop_qp_plant_kernel<32, 10, 20, KMPC_OUT_C_ROW, 1>;     // Tank_System.m (16 lanes measured 2x slower:
                                                                     // two scenarios in lock step, fewer warps)
  } else if (c.nz == 8 && c.N == 50 && !c.du_aug && c.out_mode == KMPC_OUT_C) {
    L->qp = loop_qp_plant_kernel<32, 8, 50, KMPC_OUT_C, 0>;          // BASELINE config 5
  } else {
    L->qp = loop_qp_plant_kernel<32, 0, 0, 0, 0>;
  }
  L->rls_g = 32;
  if (c.nz == 8) L->rls = loop_rls_kernel<32, 8>;
  else if (c.nz == 10) L->rls = loop_rls_kernel<32, 10>;
  else L->rls = loop_rls_kernel<32, 0>;
  const int budget = 200 * 1024;
  const int qp_ws = qp_ws_doubles(nzq, ny, c.N, identity) * (int)sizeof(double);
  const int rls_ws = rls_ws_doubles(c.nz, 2) * (int)sizeof(double);
  L->qp_spb = kLoopThreads / L->qp_g;
  while (L->qp_spb > 1 && L->qp_spb * qp_ws > budget) L->qp_spb >>= 1;
  // One scenario per block for the warp-per-scenario shapes (Tank, horizon 50): the active-set iteration
  // count of a scenario-step is heavy-tailed (Tank: 0 .. 50 factorisations, 25 % of the steps do 86 % of
  // them), and a block retires only when its slowest scenario is done -- ncu measured 10.4 active warps
  // per SM of 20 resident with 4 scenarios per block.  Single-warp blocks retire independently, and the
  // SM holds as many as its shared memory allows (Tank: 23, horizon 50: 7).
  if (L->qp_g == 32) L->qp_spb = 1;
  const int sm_smem = 225 * 1024;
  L->rls_spb = kLoopThreads / L->rls_g;
  while (L->rls_spb > 1 && L->rls_spb * rls_ws > budget) L->rls_spb >>= 1;
  L->qp_smem = L->qp_spb * qp_ws;
  L->rls_smem = L->rls_spb * rls_ws;
  return L->qp_smem <= sm_smem && L->rls_smem <= budget;
}

}  // namespace kmpc

using namespace kmpc;

namespace {
// CUDA events that are destroyed on every exit path (kmpc_closed_loop_steps_timed returns early on errors)
struct EventList {
  std::vector<cudaEvent_t> ev;
  ~EventList() {
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
  }
  cudaError_t create(size_t n) {
    ev.reserve(n);
    for (size_t i = 0; i < n; ++i) {
      cudaEvent_t e;
      const cudaError_t rc = cudaEventCreate(&e);
      if (rc != cudaSuccess) return rc;
      ev.push_back(e);
    }
    return cudaSuccess;
  }
};
}  // namespace

struct kmpc_ctx {
  LoopDev d;
  const kmpc_encoder* enc;
  int64_t step;
  int rls_started;
  LoopLaunch launch;
  bool fused;               // T steps per launch through fused_loop_kernel
  long long* d_timing;      // fused + timed: per-CTA phase cycle counters (allocated on first use)
  cudaEvent_t ready;        // recorded after the memsets of create / reset; every run waits on it, so the
                            // caller may create / reset on one stream and step on another
};

extern "C" {

int kmpc_ctx_create(kmpc_ctx** out, const kmpc_loop_config* cfg, const kmpc_loop_buffers* buf,
                    const kmpc_encoder* enc, int rls_started, void* stream) {
  if (!out || !cfg || !buf) return KMPC_ERR_ARG;
  const kmpc_loop_config& c = *cfg;
  if (c.S < 1 || c.n != 2 || c.nz < 1 || c.N < 1 || c.N > KMPC_MAX_HORIZON) return KMPC_ERR_ARG;
  const int nzq = c.nz + (c.du_aug ? 1 : 0);
  if (nzq > KMPC_MAX_NZ) return KMPC_ERR_ARG;
  if (c.out_mode < KMPC_OUT_C || c.out_mode > KMPC_OUT_C_ROW) return KMPC_ERR_ARG;
  if (c.out_mode == KMPC_OUT_C_ROW && (c.out_row < 0 || c.out_row >= c.n)) return KMPC_ERR_ARG;
  if (c.update && c.shared_model) return KMPC_ERR_ARG;  // online update needs per-scenario models
  if (c.qp_cold < 0 || c.qp_cold > 3) return KMPC_ERR_ARG;   // kmpc.h: 0 .. 3
  if (c.path != KMPC_PATH_AUTO && c.path != KMPC_PATH_GENERIC) return KMPC_ERR_ARG;
  if (!buf->x || !buf->z || !buf->u_prev || !buf->A || !buf->B || !buf->C || !buf->r ||
      !buf->params_pre || !buf->params_post)
    return KMPC_ERR_ARG;
  if (c.update && (!buf->KA || !buf->P)) return KMPC_ERR_ARG;
  if (c.update && (c.rls_flags & KMPC_RLS_UPDATE_C) && (!buf->barX || !buf->barQ)) return KMPC_ERR_ARG;
  if (c.lift_kind == KMPC_LIFTKIND_MLP) {
    if (!enc || kmpc_encoder_out_dim(enc, c.lift_mode) != c.nz) return KMPC_ERR_ARG;
  } else if (c.lift_kind == KMPC_LIFTKIND_RBF) {
    if (!buf->cx) return KMPC_ERR_ARG;
  } else {
    return KMPC_ERR_ARG;
  }
  if (!(c.lambda > 0.0)) return KMPC_ERR_ARG;
  kmpc_ctx* ctx = new (std::nothrow) kmpc_ctx();
  if (!ctx) return KMPC_ERR_ALLOC;
  ctx->d.c = c;
  ctx->d.b = *buf;
  if (ctx->d.c.max_iter <= 0) ctx->d.c.max_iter = 10 * c.N + 20;
  if (!(ctx->d.c.tol > 0.0)) ctx->d.c.tol = 1e-10;
  ctx->enc = enc;
  ctx->step = 0;
  ctx->rls_started = rls_started;
  ctx->d.z_next = nullptr;
  ctx->d.x_prev = nullptr;
  ctx->d_timing = nullptr;
  ctx->ready = nullptr;
  ctx->d.wset = nullptr;
  ctx->d.qp_x = nullptr;
  ctx->fused = fused_eligible(ctx->d.c, enc);
  if (cudaMalloc(&ctx->d.z_next, (size_t)c.S * c.nz * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&ctx->d.x_prev, (size_t)c.S * c.n * sizeof(double)) != cudaSuccess) {
    kmpc_ctx_destroy(ctx);
    return KMPC_ERR_ALLOC;
  }
  if (ctx->fused && (cudaMalloc(&ctx->d.wset, (size_t)c.S * 2 * sizeof(unsigned int)) != cudaSuccess ||
                     cudaMemsetAsync(ctx->d.wset, 0, (size_t)c.S * 2 * sizeof(unsigned int), as_stream(stream)) != cudaSuccess)) {
    kmpc_ctx_destroy(ctx);
    return KMPC_ERR_ALLOC;
  }
  {
    // generic kernels: warm start from the previous step's optimal moves (0xFF bytes = NaN = none
    // yet).  cfg.qp_cold = 1 cold-starts every QP (explicit per-context option).
    if (!ctx->fused && c.qp_cold != 1) {
      if (cudaMalloc(&ctx->d.qp_x, (size_t)c.S * c.N * sizeof(double)) != cudaSuccess ||
          cudaMemsetAsync(ctx->d.qp_x, 0xFF, (size_t)c.S * c.N * sizeof(double), as_stream(stream)) != cudaSuccess) {
        kmpc_ctx_destroy(ctx);
        return KMPC_ERR_ALLOC;
      }
    }
  }
  if (!select_launch(c, &ctx->launch)) {
    kmpc_ctx_destroy(ctx);
    return KMPC_ERR_UNSUPPORTED;
  }
  if (ensure_smem(ctx->launch.qp, ctx->launch.qp_smem) != cudaSuccess ||
      ensure_smem(ctx->launch.rls, ctx->launch.rls_smem) != cudaSuccess) {
    cudaGetLastError();
    kmpc_ctx_destroy(ctx);
    return KMPC_ERR_CUDA;
  }
  if (cudaEventCreateWithFlags(&ctx->ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventRecord(ctx->ready, as_stream(stream)) != cudaSuccess) {
    cudaGetLastError();
    kmpc_ctx_destroy(ctx);
    return KMPC_ERR_CUDA;
  }
  *out = ctx;
  return KMPC_OK;
}

int kmpc_ctx_destroy(kmpc_ctx* ctx) {
  if (!ctx) return KMPC_OK;
  if (ctx->ready) cudaEventDestroy(ctx->ready);
  if (ctx->d.z_next) cudaFree(ctx->d.z_next);
  if (ctx->d.x_prev) cudaFree(ctx->d.x_prev);
  if (ctx->d_timing) cudaFree(ctx->d_timing);
  if (ctx->d.wset) cudaFree(ctx->d.wset);
  if (ctx->d.qp_x) cudaFree(ctx->d.qp_x);
  delete ctx;
  return KMPC_OK;
}

int64_t kmpc_ctx_step_index(const kmpc_ctx* ctx) { return ctx ? ctx->step : -1; }

int kmpc_ctx_is_fused(const kmpc_ctx* ctx) { return (ctx && ctx->fused) ? 1 : 0; }

int kmpc_ctx_reset(kmpc_ctx* ctx, int rls_started, void* stream) {
  if (!ctx) return KMPC_ERR_ARG;
  ctx->step = 0;
  ctx->rls_started = rls_started;
  if (ctx->d.wset)
    KMPC_CUDA(cudaMemsetAsync(ctx->d.wset, 0, (size_t)ctx->d.c.S * 2 * sizeof(unsigned int), as_stream(stream)));
  if (ctx->d.qp_x)
    KMPC_CUDA(cudaMemsetAsync(ctx->d.qp_x, 0xFF, (size_t)ctx->d.c.S * ctx->d.c.N * sizeof(double), as_stream(stream)));
  KMPC_CUDA(cudaEventRecord(ctx->ready, as_stream(stream)));
  return KMPC_OK;
}

// One closed-loop step = qp_plant kernel -> lift kernel -> (rls kernel).  `ev` (nullable) points at
// 4 events recorded around the three launches (kmpc_closed_loop_steps_timed).
static int run_one_step(kmpc_ctx* ctx, void* stream, cudaEvent_t* ev) {
  cudaStream_t st = as_stream(stream);
  const kmpc_loop_config& c = ctx->d.c;
  const LoopLaunch& L = ctx->launch;
  const unsigned qp_grid = (unsigned)((c.S + L.qp_spb - 1) / L.qp_spb);
  const unsigned rls_grid = (unsigned)((c.S + L.rls_spb - 1) / L.rls_spb);
  const int64_t slot = (ctx->step < ctx->d.b.log_capacity) ? ctx->step : -1;
  if (ev) KMPC_CUDA(cudaEventRecord(ev[0], st));
  KMPC_CUDA(launch_pdl(L.qp, qp_grid, (unsigned)(L.qp_spb * L.qp_g), (size_t)L.qp_smem, st, ctx->d, ctx->step, slot));
  KMPC_AFTER_LAUNCH();
  if (ev) KMPC_CUDA(cudaEventRecord(ev[1], st));
  // lift(x+): into z_next when the RLS still needs the old z, else straight into z
  double* zdst = c.update ? ctx->d.z_next : ctx->d.b.z;
  int rc;
  if (c.lift_kind == KMPC_LIFTKIND_MLP)
    rc = kmpc_encode(ctx->enc, ctx->d.b.x, zdst, c.S, c.lift_mode, stream);
  else
    rc = kmpc_rbf_lift(ctx->d.b.x, ctx->d.b.cx, zdst, c.S, c.n, c.nz, c.lift_mode, stream);
  if (rc != KMPC_OK) return rc;
  if (ev) KMPC_CUDA(cudaEventRecord(ev[2], st));
  if (c.update) {
    KMPC_CUDA(launch_pdl(L.rls, rls_grid, (unsigned)(L.rls_spb * L.rls_g), (size_t)L.rls_smem, st, ctx->d,
                         ctx->rls_started ? 0 : 1));
    KMPC_AFTER_LAUNCH();
    ctx->rls_started = 1;
  }
  if (ev) KMPC_CUDA(cudaEventRecord(ev[3], st));
  ctx->step += 1;
  return KMPC_OK;
}

static int run_fused(kmpc_ctx* ctx, int T, long long* timing, int* grid, void* stream) {
  if (T == 0) return KMPC_OK;
  const int first = (ctx->d.c.update && !ctx->rls_started) ? 1 : 0;
  const int rc = fused_launch(ctx->d, ctx->enc, ctx->step, T, first, timing, grid, as_stream(stream));
  if (rc != KMPC_OK) return rc;
  if (ctx->d.c.update) ctx->rls_started = 1;
  ctx->step += T;
  return KMPC_OK;
}

int kmpc_closed_loop_steps(kmpc_ctx* ctx, int T, void* stream) {
  if (!ctx || T < 0) return KMPC_ERR_ARG;
  KMPC_CUDA(cudaStreamWaitEvent(as_stream(stream), ctx->ready, 0));
  if (ctx->fused) return run_fused(ctx, T, nullptr, nullptr, stream);
  for (int t = 0; t < T; ++t) {
    const int rc = run_one_step(ctx, stream, nullptr);
    if (rc != KMPC_OK) return rc;
  }
  return KMPC_OK;
}

int kmpc_closed_loop_steps_timed(kmpc_ctx* ctx, int T, void* stream, float* ms) {
  if (!ctx || T < 1 || !ms) return KMPC_ERR_ARG;
  KMPC_CUDA(cudaStreamWaitEvent(as_stream(stream), ctx->ready, 0));
  if (ctx->fused) {
    // one launch; the kernel accumulates clock64() per phase in every CTA, the launch itself is
    // bracketed by CUDA events: ms[k] = launch time x mean over CTAs of the phase's cycle share
    const int kMaxCtas = 4096;
    if (!ctx->d_timing && cudaMalloc(&ctx->d_timing, sizeof(long long) * 4 * kMaxCtas) != cudaSuccess)
      return KMPC_ERR_ALLOC;
    cudaStream_t st = as_stream(stream);
    EventList el;
    KMPC_CUDA(el.create(2));
    const cudaEvent_t e0 = el.ev[0], e1 = el.ev[1];
    KMPC_CUDA(cudaEventRecord(e0, st));
    int grid = 0;
    int rc = run_fused(ctx, T, ctx->d_timing, &grid, stream);
    if (rc == KMPC_OK && cudaEventRecord(e1, st) != cudaSuccess) rc = KMPC_ERR_CUDA;
    if (rc == KMPC_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = KMPC_ERR_CUDA;
    ms[0] = ms[1] = ms[2] = 0.f;
    if (rc == KMPC_OK) {
      float total = 0.f;
      cudaEventElapsedTime(&total, e0, e1);
      std::vector<long long> h((size_t)4 * grid);
      if (cudaMemcpy(h.data(), ctx->d_timing, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = KMPC_ERR_CUDA;
      double share[3] = {0, 0, 0};
      for (int i = 0; rc == KMPC_OK && i < grid; ++i)
        for (int k = 0; k < 3; ++k) share[k] += (double)h[4 * i + k] / (double)(h[4 * i + 3] > 0 ? h[4 * i + 3] : 1);
      for (int k = 0; k < 3; ++k) ms[k] = (float)(total * share[k] / (grid > 0 ? grid : 1));
#ifdef KMPC_PROFILING
      if (rc == KMPC_OK && getenv("KMPC_DEBUG_TIMING")) {
        long long mx = 0, mn = h[3];
        double mean = 0.0;
        for (int i = 0; i < grid; ++i) {
          mx = h[4 * i + 3] > mx ? h[4 * i + 3] : mx;
          mn = h[4 * i + 3] < mn ? h[4 * i + 3] : mn;
          mean += (double)h[4 * i + 3] / grid;
        }
        // cycles of quarter 0 of every CTA: the launch lasts as long as its slowest tile
        fprintf(stderr, "[kmpc] fused timed: grid %d, T %d, %.3f ms, CTA cycles min %lld mean %.0f max %lld -> SM clock %.0f MHz\n",
                grid, T, total, mn, mean, mx, mx / (total * 1e3));
      }
#endif
    }
    return rc;
  }
  if (T > 1024) return KMPC_ERR_ARG;
  EventList el;
  KMPC_CUDA(el.create((size_t)4 * T));
  std::vector<cudaEvent_t>& ev = el.ev;
  int rc = KMPC_OK;
  for (int t = 0; t < T && rc == KMPC_OK; ++t) rc = run_one_step(ctx, stream, ev.data() + 4 * t);
  // synchronise even after a failed step: the events recorded so far must not outlive their owner
  if (cudaStreamSynchronize(as_stream(stream)) != cudaSuccess && rc == KMPC_OK) rc = KMPC_ERR_CUDA;
  ms[0] = ms[1] = ms[2] = 0.f;
  if (rc == KMPC_OK) {
    for (int t = 0; t < T; ++t)
      for (int k = 0; k < 3; ++k) {
        float v = 0.f;
        cudaEventElapsedTime(&v, ev[4 * t + k], ev[4 * t + k + 1]);
        ms[k] += v;
      }
  }
  return rc;
}

}  // extern "C"
