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

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
loop_qp_plant_kernel(LoopDev d, int64_t step, int64_t log_slot) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5;
  const int64_t s = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (s >= d.c.S) return;
  const bool identity = d.c.out_mode == KMPC_OUT_IDENTITY;
  loop_qp_plant_scenario(d, s, step, log_slot,
                         smem + (size_t)warp * qp_ws_doubles(loop_nzq(d.c), loop_ny(d.c), d.c.N, identity));
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
loop_rls_kernel(LoopDev d, int first) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5;
  const int64_t s = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (s >= d.c.S) return;
  loop_rls_scenario(d, s, first, smem + (size_t)warp * rls_ws_doubles(d.c.nz, d.c.n));
}

}  // namespace kmpc

using namespace kmpc;

struct kmpc_ctx {
  LoopDev d;
  const kmpc_encoder* enc;
  int64_t step;
  int rls_started;
  int qp_smem, rls_smem;
};

extern "C" {

int kmpc_ctx_create(kmpc_ctx** out, const kmpc_loop_config* cfg, const kmpc_loop_buffers* buf,
                    const kmpc_encoder* enc, int rls_started, void* stream) {
  (void)stream;
  if (!out || !cfg || !buf) return KMPC_ERR_ARG;
  const kmpc_loop_config& c = *cfg;
  if (c.S < 1 || c.n != 2 || c.nz < 1 || c.N < 1 || c.N > KMPC_MAX_HORIZON) return KMPC_ERR_ARG;
  const int nzq = c.nz + (c.du_aug ? 1 : 0);
  if (nzq > KMPC_MAX_NZ) return KMPC_ERR_ARG;
  if (c.out_mode < KMPC_OUT_C || c.out_mode > KMPC_OUT_C_ROW) return KMPC_ERR_ARG;
  if (c.out_mode == KMPC_OUT_C_ROW && (c.out_row < 0 || c.out_row >= c.n)) return KMPC_ERR_ARG;
  if (c.update && c.shared_model) return KMPC_ERR_ARG;  // online update needs per-scenario models
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
  const bool identity = c.out_mode == KMPC_OUT_IDENTITY;
  const int ny = identity ? nzq : (c.out_mode == KMPC_OUT_C ? c.n : 1);
  ctx->qp_smem = kWarpsPerBlock * qp_ws_doubles(nzq, ny, c.N, identity) * (int)sizeof(double);
  ctx->rls_smem = kWarpsPerBlock * rls_ws_doubles(c.nz, c.n) * (int)sizeof(double);
  ctx->d.z_next = nullptr;
  ctx->d.x_prev = nullptr;
  if (cudaMalloc(&ctx->d.z_next, (size_t)c.S * c.nz * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&ctx->d.x_prev, (size_t)c.S * c.n * sizeof(double)) != cudaSuccess) {
    kmpc_ctx_destroy(ctx);
    return KMPC_ERR_ALLOC;
  }
  if (ensure_smem(loop_qp_plant_kernel, ctx->qp_smem) != cudaSuccess ||
      ensure_smem(loop_rls_kernel, ctx->rls_smem) != cudaSuccess) {
    kmpc_ctx_destroy(ctx);
    return KMPC_ERR_CUDA;
  }
  *out = ctx;
  return KMPC_OK;
}

int kmpc_ctx_destroy(kmpc_ctx* ctx) {
  if (!ctx) return KMPC_OK;
  if (ctx->d.z_next) cudaFree(ctx->d.z_next);
  if (ctx->d.x_prev) cudaFree(ctx->d.x_prev);
  delete ctx;
  return KMPC_OK;
}

int64_t kmpc_ctx_step_index(const kmpc_ctx* ctx) { return ctx ? ctx->step : -1; }

// One closed-loop step = qp_plant kernel -> lift kernel -> (rls kernel).  `ev` (nullable) points at
// 4 events recorded around the three launches (kmpc_closed_loop_steps_timed).
static int run_one_step(kmpc_ctx* ctx, void* stream, cudaEvent_t* ev) {
  cudaStream_t st = as_stream(stream);
  const kmpc_loop_config& c = ctx->d.c;
  const unsigned grid = (unsigned)((c.S + kWarpsPerBlock - 1) / kWarpsPerBlock);
  const int64_t slot = (ctx->step < ctx->d.b.log_capacity) ? ctx->step : -1;
  if (ev) KMPC_CUDA(cudaEventRecord(ev[0], st));
  loop_qp_plant_kernel<<<grid, kWarpsPerBlock * 32, ctx->qp_smem, st>>>(ctx->d, ctx->step, slot);
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
    loop_rls_kernel<<<grid, kWarpsPerBlock * 32, ctx->rls_smem, st>>>(ctx->d, ctx->rls_started ? 0 : 1);
    KMPC_AFTER_LAUNCH();
    ctx->rls_started = 1;
  }
  if (ev) KMPC_CUDA(cudaEventRecord(ev[3], st));
  ctx->step += 1;
  return KMPC_OK;
}

int kmpc_closed_loop_steps(kmpc_ctx* ctx, int T, void* stream) {
  if (!ctx || T < 0) return KMPC_ERR_ARG;
  for (int t = 0; t < T; ++t) {
    const int rc = run_one_step(ctx, stream, nullptr);
    if (rc != KMPC_OK) return rc;
  }
  return KMPC_OK;
}

int kmpc_closed_loop_steps_timed(kmpc_ctx* ctx, int T, void* stream, float* ms) {
  if (!ctx || T < 1 || T > 1024 || !ms) return KMPC_ERR_ARG;
  std::vector<cudaEvent_t> ev((size_t)4 * T);
  for (auto& e : ev) KMPC_CUDA(cudaEventCreate(&e));
  int rc = KMPC_OK;
  for (int t = 0; t < T && rc == KMPC_OK; ++t) rc = run_one_step(ctx, stream, ev.data() + 4 * t);
  if (rc == KMPC_OK && cudaStreamSynchronize(as_stream(stream)) != cudaSuccess) rc = KMPC_ERR_CUDA;
  ms[0] = ms[1] = ms[2] = 0.f;
  if (rc == KMPC_OK) {
    for (int t = 0; t < T; ++t)
      for (int k = 0; k < 3; ++k) {
        float v = 0.f;
        cudaEventElapsedTime(&v, ev[4 * t + k], ev[4 * t + k + 1]);
        ms[k] += v;
      }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

}  // extern "C"
