// lift.cu -- theta_E encoder MLP (fp64, CUDA-core register-tiled batched GEMM chain).
//
// Reference: duffing.py:21-29 (`nn.Sequential(Linear(2,100), ReLU, Linear(100,100), ReLU,
// Linear(100,100), ReLU, Linear(100,8))`), Encoder_Tank.m:3-5 (3 layers, nz = 10).
//
// Layout: one CTA lifts tiles of kTileS = 32 scenarios through ALL layers; activations stay in
// shared memory, k-major ([k][scenario]) so a thread's 4 consecutive scenarios are one 32-byte
// read that is broadcast to the 4 lanes sharing the row group.  Weights are stored transposed and
// padded ([in][out_pad], out_pad multiple of 4), packed layer after layer in one device buffer.
// Each thread owns a 4 (scenarios) x 4 (outputs) register tile: 16 DFMA per 8 x 8-byte operands.
//
// encoder_units_kernel (default): the whole packed weight set (175 KB for 2-100-100-100-8) is
// brought into shared memory ONCE per CTA by the TMA engine (cp.async.bulk, one mbarrier per
// layer); CTAs are persistent; the CTA is four independent lift units of 2 warps x 8 rows running
// encoder.cuh: lift_unit<2> (ping-pong activations, one unit-local barrier per layer, software-
// pipelined DMMA k-loop, split-K last layer).  The layer GEMMs run on the fp64 tensor path
// (mma.sync m8n8k4 f64 -- tcgen05 has no f64 kind, and DMMA measures the same 37 TFLOP/s as DFMA on
// B200, but needs 8x fewer shared-memory wavefronts per MAC, which is what bounded the CUDA-core
// version).  Row strides of the weight arrays (== 4 or 12 mod 16) and the activation pitch (8) make
// every fragment load bank-conflict free.
// encoder_mma_kernel (nets with more than 16 outputs): same staging, CTA-wide 32-row tiles.
// encoder_kernel (fallback for nets that do not fit 227 KB): CUDA-core 4x4 register tiles,
// weights read from global/L2.
#include "encoder.cuh"

namespace kmpc {

constexpr int kEncWarps = 7;       // 7 warps x 4 column groups x 4 outputs = 112 outputs per pass
constexpr int kEncThreads = kEncWarps * 32;

// One layer of the 4x4 register-tiled GEMM for this thread: acc = bias + act_in^T W.
// WT_LD: functor-free: weights read with plain loads from `wt` (shared or global pointer).
template <bool kGlobalWeights>
__device__ __forceinline__ void enc_layer_tile(const double* __restrict__ hp, const double* __restrict__ wp,
                                               const double* __restrict__ bias4, int in, int outp,
                                               double (&acc)[4][4]) {
  double2 b01, b23;
  if (kGlobalWeights) {
    b01 = __ldg(reinterpret_cast<const double2*>(bias4));
    b23 = __ldg(reinterpret_cast<const double2*>(bias4 + 2));
  } else {
    b01 = *reinterpret_cast<const double2*>(bias4);
    b23 = *reinterpret_cast<const double2*>(bias4 + 2);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    acc[r][0] = b01.x;
    acc[r][1] = b01.y;
    acc[r][2] = b23.x;
    acc[r][3] = b23.y;
  }
#pragma unroll 4
  for (int k = 0; k < in; ++k) {
    const double2 h01 = *reinterpret_cast<const double2*>(hp + k * kTileS);
    const double2 h23 = *reinterpret_cast<const double2*>(hp + k * kTileS + 2);
    double2 w01, w23;
    if (kGlobalWeights) {
      w01 = __ldg(reinterpret_cast<const double2*>(wp + (size_t)k * outp));
      w23 = __ldg(reinterpret_cast<const double2*>(wp + (size_t)k * outp + 2));
    } else {
      w01 = *reinterpret_cast<const double2*>(wp + k * outp);
      w23 = *reinterpret_cast<const double2*>(wp + k * outp + 2);
    }
    const double h[4] = {h01.x, h01.y, h23.x, h23.y};
    const double w[4] = {w01.x, w01.y, w23.x, w23.y};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = fma(h[r], w[c], acc[r][c]);
  }
}

// ReLU + store to the next activation buffer, or final store to global z.
__device__ __forceinline__ void enc_store_tile(const double (&acc)[4][4], bool last, double* act_out, int cg,
                                               int rg, double* __restrict__ z, int64_t row0, int64_t S,
                                               int out, int out_dim, int off, int lift_mode,
                                               const double* __restrict__ z0) {
  if (!last) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double2 o01, o23;
      o01.x = (acc[0][c] < 0.0 ? 0.0 : acc[0][c]);
      o01.y = (acc[1][c] < 0.0 ? 0.0 : acc[1][c]);
      o23.x = (acc[2][c] < 0.0 ? 0.0 : acc[2][c]);
      o23.y = (acc[3][c] < 0.0 ? 0.0 : acc[3][c]);
      double* op = act_out + (4 * cg + c) * kTileS + 4 * rg;
      *reinterpret_cast<double2*>(op) = o01;
      *reinterpret_cast<double2*>(op + 2) = o23;
    }
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int64_t row = row0 + 4 * rg + r;
      if (row < S) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col = 4 * cg + c;
          if (col < out) {
            double v = acc[r][c];
            if (lift_mode != KMPC_LIFT_RAW) v -= z0[col];
            z[row * out_dim + off + col] = v;
          }
        }
      }
    }
  }
}

// Persistent CTAs, weights resident in shared memory (loaded once by TMA bulk copies), layer GEMMs
// on the fp64 tensor path (encoder.cuh: encoder_layers).
__global__ void __launch_bounds__(kMmaThreads, 1)
encoder_mma_kernel(EncParams p, const double* __restrict__ x, double* __restrict__ z, int64_t S,
                   int lift_mode, int out_dim, int64_t num_tiles) {
  extern __shared__ __align__(16) double smem[];
  double* act = smem;
  double* wsm = smem + p.actw * kActStride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + p.total_w);
  const int tid = threadIdx.x;
  if (tid == 0) encoder_weights_init_barriers(p, bars);
  __syncthreads();
  if (tid == 0) encoder_weights_issue(p, wsm, bars);
  // the weight copies above do not depend on the previous kernel: only now wait for it (PDL)
  pdl_wait();
  pdl_launch_dependents();
  const int n = p.dims[0];
  const int off = (lift_mode == KMPC_LIFT_STACK) ? n : 0;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kTileS;
    // layer-0 input, k-major, rows [n, inpad) zero
    for (int e = tid; e < kTileS * p.inpad[0]; e += kMmaThreads) {
      const int k = e / kTileS, r = e - k * kTileS;
      double v = 0.0;
      if (k < n && row0 + r < S) {
        v = x[(row0 + r) * n + k];
        if (lift_mode == KMPC_LIFT_STACK) z[(row0 + r) * out_dim + k] = v;
      }
      act[k * kActStride + r] = v;
    }
    __syncthreads();
    encoder_layers(p, act, act, wsm, bars, [&](int r, int col, double v) {
      const int64_t row = row0 + r;
      if (row < S) {
        if (lift_mode != KMPC_LIFT_RAW) v -= p.z0[col];
        z[row * out_dim + off + col] = v;
      }
    });
  }
}

// Default kernel: the CTA is kEncUnits independent lift units of 2 warps x 8 rows (encoder.cuh:
// lift_unit<2>, the same code the fused closed-loop kernel runs) with IN-PLACE activations: 6.5 KB of
// shared memory per unit, so eight units (16 warps, four per scheduler) fit beside the weights.  The
// units drift apart; with four warps per scheduler the tensor pipe keeps issuing while other units
// sit at their barriers or in their epilogues (two warps per scheduler: 66 % DMMA issue rate).
constexpr int kEncUnits = 8;
constexpr int kEncUnitThreads = kEncUnits * 64;
struct UnitsSmem {  // offsets in doubles
  int region, in0, yout, ypitch, wsm, bars, total_bytes;
};
inline UnitsSmem units_smem_layout(const EncParams& p) {
  UnitsSmem L;
  const int out = p.dims[p.n_layers];
  const int actbuf = p.actw * kActPitch;
  L.ypitch = (out + 1) & ~1;
  L.yout = 128;                                  // beyond the split-K partials (2 warps x 64)
  L.in0 = actbuf - 4 * kUnitRows;                // the pad k-rows of the activation buffer
  L.region = actbuf;
  L.wsm = kEncUnits * L.region;
  L.bars = (L.wsm + p.total_w + 1) & ~1;
  L.total_bytes = (L.bars + KMPC_MAX_LAYERS) * 8;
  return L;
}
inline bool units_eligible(const EncParams& p, int max_smem) {
  const int out = p.dims[p.n_layers];
  if (p.n_layers < 2 || out > 16 || p.dims[0] > 4) return false;
  int widest = 0;
  for (int l = 1; l < p.n_layers; ++l) widest = p.dims[l] > widest ? p.dims[l] : widest;
  // in0 lives in k-rows [actw - 4, actw): they must be pad rows of every hidden layer's OUTPUT (their
  // weights rows are zero), and y (8 x ypitch from 128) must end below them
  if (p.actw - 4 < widest) return false;
  if (128 + kUnitRows * ((out + 1) & ~1) > (p.actw - 4) * kActPitch) return false;
  return units_smem_layout(p).total_bytes <= max_smem;
}

__global__ void __launch_bounds__(kEncUnitThreads, 1)
encoder_units_kernel(EncParams p, UnitsSmem L, const double* __restrict__ x, double* __restrict__ z,
                     int64_t S, int lift_mode, int out_dim, int64_t n_blocks) {
  extern __shared__ __align__(16) double smem[];
  double* wsm = smem + L.wsm;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  const int tid = threadIdx.x;
  if (tid == 0) encoder_weights_init_barriers(p, bars);
  __syncthreads();
  if (tid == 0) encoder_weights_issue(p, wsm, bars);
  pdl_wait();   // the weight copies above do not depend on the previous kernel: only now wait for it
  pdl_launch_dependents();
  encoder_weights_wait_all(p, bars);
  const int unit = tid >> 6, wl = (tid >> 5) & 1, t64 = tid & 63, bar = 1 + unit;
  double* region = smem + unit * L.region;
  double* in0 = region + L.in0;
  double* yo = region + L.yout;
  const int n = p.dims[0], out = p.dims[p.n_layers];
  const int off = (lift_mode == KMPC_LIFT_STACK) ? n : 0;
  for (int64_t rb = (int64_t)blockIdx.x * kEncUnits + unit; rb < n_blocks; rb += (int64_t)gridDim.x * kEncUnits) {
    const int64_t row0 = rb * kUnitRows;
    if (t64 < 4 * kUnitRows) {   // layer-0 input, k-major, rows [n, 4) zero
      const int k = t64 >> 3, r = t64 & 7;
      double v = 0.0;
      if (k < n && row0 + r < S) {
        v = x[(row0 + r) * n + k];
        if (lift_mode == KMPC_LIFT_STACK) z[(row0 + r) * out_dim + k] = v;
      }
      in0[act_index(k, r)] = v;
    }
    group_barrier<64>(bar);
    // units u, u + 2, .. share two schedulers: alternate the warp that owns the odd n-tile
    lift_unit<2>(p, in0, region, region, yo, L.ypitch, wsm, wl ^ ((unit >> 1) & 1), tid & 31, bar);
    for (int e = t64; e < kUnitRows * out; e += 64) {
      const int r = e / out, c = e - r * out;
      if (row0 + r < S) {
        double v = yo[r * L.ypitch + c];
        if (lift_mode != KMPC_LIFT_RAW) v -= p.z0[c];
        z[(row0 + r) * out_dim + off + c] = v;
      }
    }
    group_barrier<64>(bar);   // the next block's layers overwrite the outputs
  }
}

__global__ void __launch_bounds__(kEncThreads)
encoder_kernel(EncParams p, const double* __restrict__ x, double* __restrict__ z, int64_t S,
               int lift_mode, int out_dim) {
  extern __shared__ __align__(16) double smem[];
  double* act_in = smem;
  double* act_out = smem + KMPC_MAX_WIDTH * kTileS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * kTileS;
  const int n = p.dims[0];
  for (int e = tid; e < kTileS * n; e += kEncThreads) {
    const int r = e / n, k = e - r * n;
    const double v = (row0 + r < S) ? x[(row0 + r) * n + k] : 0.0;
    act_in[k * kTileS + r] = v;
    if (lift_mode == KMPC_LIFT_STACK && row0 + r < S) z[(row0 + r) * out_dim + k] = v;
  }
  __syncthreads();
  const int rg = lane & 7, cgl = lane >> 3;
  const int off = (lift_mode == KMPC_LIFT_STACK) ? n : 0;
  for (int l = 0; l < p.n_layers; ++l) {
    const int in = p.dims[l], outp = p.pad[l + 1], out = p.dims[l + 1];
    const int ncg = outp >> 2;
    const bool last = (l == p.n_layers - 1);
    for (int cg0 = 0; cg0 < ncg; cg0 += kEncWarps * 4) {
      const int cg = cg0 + warp * 4 + cgl;
      if (cg < ncg) {
        double acc[4][4];
        enc_layer_tile<true>(act_in + 4 * rg, p.wt[l] + 4 * cg, p.b[l] + 4 * cg, in, outp, acc);
        enc_store_tile(acc, last, act_out, cg, rg, z, row0, S, out, out_dim, off, lift_mode, p.z0);
      }
    }
    __syncthreads();
    double* t = act_in;
    act_in = act_out;
    act_out = t;
  }
}

}  // namespace kmpc

namespace kmpc {
int tc_encode_launch(const kmpc_encoder* enc, const double* x, double* z, int64_t S, int lift_mode,
                     cudaStream_t st);   // tc_lift.cu
int gram_accumulate_impl(const double* psi, const double* psi_next, const double* u, const double* x,
                         int64_t M, int nz, int n, double* pack, int seg, void* stream);   // edmd.cu

// rows of a chunk of trajectories for the trajectory-aware Gram: trajectory t contributes its
// n_step snapshot states x and the successor y of its last snapshot, n_step + 1 consecutive rows
__global__ void traj_rows_kernel(const double* __restrict__ x, const double* __restrict__ y, int n,
                                 int64_t n_traj, int n_step, double* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_traj * (n_step + 1)) return;
  const int64_t t = r / (n_step + 1);
  const int j = (int)(r - t * (n_step + 1));
  const double* src = (j < n_step) ? x + (t * n_step + j) * n : y + (t * n_step + n_step - 1) * n;
  for (int k = 0; k < n; ++k) out[r * n + k] = src[k];
}
}  // namespace kmpc

using namespace kmpc;


static int launch_encoder(const kmpc_encoder* enc, const double* x, double* z, int64_t S,
                          int lift_mode, cudaStream_t st) {
  const int64_t tiles = (S + kTileS - 1) / kTileS;
  if (tiles > 0x7fffffff) return KMPC_ERR_ARG;
  const int out_dim = kmpc_encoder_out_dim(enc, lift_mode);
#ifdef KMPC_PROFILING
  static const bool units_off = [] {   // profiling builds only: KMPC_ENC_UNITS=0 selects the CTA-wide kernel
    const char* e = getenv("KMPC_ENC_UNITS");
    return e && e[0] == '0';
  }();
#else
  constexpr bool units_off = false;
#endif
  if (!units_off && units_eligible(enc->p, enc->max_smem_optin)) {
    const UnitsSmem L = units_smem_layout(enc->p);
    KMPC_CUDA(ensure_smem(encoder_units_kernel, L.total_bytes));
    const int64_t blocks = (S + kUnitRows - 1) / kUnitRows, ctas = (blocks + kEncUnits - 1) / kEncUnits;
    const unsigned grid = (unsigned)(ctas < enc->num_sms ? ctas : enc->num_sms);
    KMPC_CUDA(launch_pdl(encoder_units_kernel, grid, (unsigned)kEncUnitThreads, (size_t)L.total_bytes, st, enc->p, L, x, z,
                         S, lift_mode, out_dim, blocks));
  } else if (enc->smem_bytes > 0) {
    KMPC_CUDA(ensure_smem(encoder_mma_kernel, enc->smem_bytes));
    const unsigned grid = (unsigned)(tiles < enc->num_sms ? tiles : enc->num_sms);
    KMPC_CUDA(launch_pdl(encoder_mma_kernel, grid, (unsigned)kMmaThreads, (size_t)enc->smem_bytes, st, enc->p, x, z,
                         S, lift_mode, out_dim, tiles));
  } else {
    const int smem = 2 * KMPC_MAX_WIDTH * kTileS * (int)sizeof(double);
    KMPC_CUDA(ensure_smem(encoder_kernel, smem));
    encoder_kernel<<<(unsigned)tiles, kEncThreads, smem, st>>>(enc->p, x, z, S, lift_mode, out_dim);
  }
  KMPC_AFTER_LAUNCH();
  return KMPC_OK;
}

extern "C" {

int kmpc_encoder_create(kmpc_encoder** out, const double* const* W, const double* const* b,
                        const int* dims, int n_layers, void* stream) {
  if (!out || !W || !b || !dims || n_layers < 1 || n_layers > KMPC_MAX_LAYERS) return KMPC_ERR_ARG;
  for (int l = 0; l <= n_layers; ++l)
    if (dims[l] < 1 || dims[l] > KMPC_MAX_WIDTH) return KMPC_ERR_ARG;
  if (dims[0] > 16) return KMPC_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  kmpc_encoder* enc = new kmpc_encoder();
  enc->p.n_layers = n_layers;
  for (int l = 0; l <= n_layers; ++l) {
    enc->p.dims[l] = dims[l];
    enc->p.pad[l] = (dims[l] + 3) & ~3;
  }
  auto fail = [&](int code) {
    kmpc_encoder_destroy(enc);
    return code;
  };
  // (a) tensor-path layout, exactly what encoder_mma_kernel keeps in shared memory:
  //     per layer W^T as [inpad][wstride] (zero rows/cols beyond in/out; wstride == 4 mod 16 so the
  //     4 k-rows of a B fragment hit disjoint banks) followed by the bias padded to 8.
  // (b) fallback layout for encoder_kernel: W^T as [in][pad4(out)], bias pad4(out).
  std::vector<double> packed, flat;
  int actw = 4;
  std::vector<size_t> foff(n_layers);
  for (int l = 0; l < n_layers; ++l) {
    const int in = dims[l], o = dims[l + 1], op = enc->p.pad[l + 1];
    const int inpad = (in + 3) & ~3, o8 = (o + 7) & ~7;
    int wsd = (o + 3) & ~3;
    while ((wsd & 7) != 4) wsd += 4;   // == 4 or 12 (mod 16): both conflict free
    enc->p.inpad[l] = inpad;
    enc->p.wstride[l] = wsd;
    enc->p.woff[l] = (int)packed.size();
    enc->p.wlen[l] = inpad * wsd + o8;
    packed.resize(packed.size() + (size_t)inpad * wsd + o8, 0.0);
    double* wt = packed.data() + enc->p.woff[l];
    double* bp = wt + (size_t)inpad * wsd;
    foff[l] = flat.size();
    flat.resize(flat.size() + (size_t)in * op + op, 0.0);
    double* fw = flat.data() + foff[l];
    double* fb = fw + (size_t)in * op;
    for (int i = 0; i < o; ++i) {
      bp[i] = b[l][i];
      fb[i] = b[l][i];
      for (int k = 0; k < in; ++k) {
        wt[(size_t)k * wsd + i] = W[l][(size_t)i * in + k];
        fw[(size_t)k * op + i] = W[l][(size_t)i * in + k];
      }
    }
    if (inpad > actw) actw = inpad;
    if (l + 1 < n_layers && ((o + 7) & ~7) > actw) actw = (o + 7) & ~7;
  }
  packed.resize(packed.size() + 16, 0.0);  // slack: the last n-tile of a layer may read past its row
  enc->p.total_w = (int)packed.size();
  enc->p.actw = actw;
  double *dpk = nullptr, *dfl = nullptr;
  if (cudaMalloc(&dpk, packed.size() * sizeof(double)) != cudaSuccess) return fail(KMPC_ERR_ALLOC);
  enc->owned.push_back(dpk);
  if (cudaMalloc(&dfl, flat.size() * sizeof(double)) != cudaSuccess) return fail(KMPC_ERR_ALLOC);
  enc->owned.push_back(dfl);
  // pageable-source async copies complete w.r.t. the host before returning
  if (cudaMemcpyAsync(dpk, packed.data(), packed.size() * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(dfl, flat.data(), flat.size() * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess)
    return fail(KMPC_ERR_CUDA);
  enc->p.packed = dpk;
  for (int l = 0; l < n_layers; ++l) {
    enc->p.wt[l] = dfl + foff[l];
    enc->p.b[l] = dfl + foff[l] + (size_t)dims[l] * enc->p.pad[l + 1];
  }
  {
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&enc->num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    enc->max_smem_optin = max_smem;
    const size_t need = ((size_t)actw * kActStride + packed.size()) * sizeof(double) + KMPC_MAX_LAYERS * 8;
    enc->smem_bytes = (need <= (size_t)max_smem) ? (int)need : 0;
  }
  // theta(0) for the OFFSET / STACK lift modes (Koopman_update.m:67)
  double* scratch = nullptr;
  if (cudaMalloc(&scratch, (16 + KMPC_MAX_WIDTH) * sizeof(double)) != cudaSuccess) return fail(KMPC_ERR_ALLOC);
  enc->owned.push_back(scratch);
  enc->d_z0 = scratch + 16;
  enc->p.z0 = enc->d_z0;
  if (cudaMemsetAsync(scratch, 0, (16 + KMPC_MAX_WIDTH) * sizeof(double), st) != cudaSuccess)
    return fail(KMPC_ERR_CUDA);
  int rc = launch_encoder(enc, scratch, enc->d_z0, 1, KMPC_LIFT_RAW, st);
  if (rc != KMPC_OK) return fail(rc);
  if (cudaStreamSynchronize(st) != cudaSuccess) return fail(KMPC_ERR_CUDA);
  // split-precision image for the tcgen05 lift (KMPC_PREC_TC); null when the net does not fit it
  enc->tc = tc_state_create(W, b, dims, n_layers, enc->d_z0, st);
  if (cudaEventCreateWithFlags(&enc->ws_done, cudaEventDisableTiming) != cudaSuccess) return fail(KMPC_ERR_CUDA);
  *out = enc;
  return KMPC_OK;
}

int kmpc_encoder_destroy(kmpc_encoder* enc) {
  if (!enc) return KMPC_OK;
  for (double* p : enc->owned) cudaFree(p);
  if (enc->d_ws) cudaFree(enc->d_ws);
  if (enc->ws_done) cudaEventDestroy(enc->ws_done);
  tc_state_destroy(enc->tc);
  delete enc;
  return KMPC_OK;
}

int kmpc_encoder_out_dim(const kmpc_encoder* enc, int lift_mode) {
  if (!enc) return KMPC_ERR_ARG;
  const int nz = enc->p.dims[enc->p.n_layers];
  return lift_mode == KMPC_LIFT_STACK ? nz + enc->p.dims[0] : nz;
}

static int encode_any(const kmpc_encoder* enc, const double* x, double* z, int64_t S, int lift_mode,
                      int precision, cudaStream_t st) {
  if (precision == KMPC_PREC_TC) return tc_encode_launch(enc, x, z, S, lift_mode, st);
  return launch_encoder(enc, x, z, S, lift_mode, st);
}

int kmpc_encode_ex(const kmpc_encoder* enc, const double* x, double* z, int64_t S, int lift_mode,
                   int precision, void* stream) {
  if (!enc || S < 0) return KMPC_ERR_ARG;
  if (lift_mode < KMPC_LIFT_RAW || lift_mode > KMPC_LIFT_STACK) return KMPC_ERR_ARG;
  if (precision != KMPC_PREC_FP64 && precision != KMPC_PREC_TC) return KMPC_ERR_ARG;
  if (precision == KMPC_PREC_TC && !enc->tc) return KMPC_ERR_UNSUPPORTED;
  if (S == 0) return KMPC_OK;
  if (!x || !z) return KMPC_ERR_ARG;
  return encode_any(enc, x, z, S, lift_mode, precision, as_stream(stream));
}

int kmpc_encode(const kmpc_encoder* enc, const double* x, double* z, int64_t S, int lift_mode,
                void* stream) {
  return kmpc_encode_ex(enc, x, z, S, lift_mode, KMPC_PREC_FP64, stream);
}

int kmpc_encoder_has_tc(const kmpc_encoder* enc) { return (enc && enc->tc) ? 1 : 0; }

// The lift workspace of a handle: sized from the actual widths, serialised across callers.
namespace {
// Rows lifted per chunk.  Round 1 used 2^18 (the lifted chunk stays in L2), but at the tcgen05 lift's
// 3.4e9 rows/s a 10 M-snapshot regression then spent 40 % of its time in the 39 x 3 short launches and
// their atomics tails (profiles/r2/launches_edmd_tc_chunk18.txt); 2^22 rows (256 MiB of lifted states,
// one HBM round trip at 6.5 TB/s = 0.1 ms) leaves 3 chunks.
constexpr int64_t kGramChunk = 1 << 22;
struct WsLock {
  kmpc_encoder* enc;
  cudaStream_t st;
  std::unique_lock<std::mutex> lk;
  WsLock(kmpc_encoder* e, cudaStream_t s) : enc(e), st(s), lk(e->ws_mu) {}
  // device-side order against the previous user of the workspace (possibly another stream)
  cudaError_t acquire(size_t doubles) {
    cudaError_t rc = cudaStreamWaitEvent(st, enc->ws_done, 0);
    if (rc != cudaSuccess) return rc;
    if (enc->ws_doubles < doubles) {
      if (enc->d_ws) {
        rc = cudaEventSynchronize(enc->ws_done);   // nobody may still be reading the old block
        if (rc != cudaSuccess) return rc;
        cudaFree(enc->d_ws);
        enc->d_ws = nullptr;
        enc->ws_doubles = 0;
      }
      rc = cudaMalloc(&enc->d_ws, doubles * sizeof(double));
      if (rc != cudaSuccess) return rc;
      enc->ws_doubles = doubles;
    }
    return cudaSuccess;
  }
  ~WsLock() { cudaEventRecord(enc->ws_done, st); }
};
}  // namespace

// fused lift + Gram: lifts chunks of snapshots into an L2-sized workspace owned by the encoder
// handle and accumulates the Gram pack from it, so PHIX / PHIY never round-trip HBM in full.
int kmpc_gram_from_snapshots_ex(const kmpc_encoder* enc_c, int lift_mode, int precision, const double* x,
                                const double* y, const double* u, int64_t M, double* pack,
                                void* stream) {
  if (!enc_c || !x || !y || !u || !pack || M < 0) return KMPC_ERR_ARG;
  if (precision != KMPC_PREC_FP64 && precision != KMPC_PREC_TC) return KMPC_ERR_ARG;
  kmpc_encoder* enc = const_cast<kmpc_encoder*>(enc_c);
  if (precision == KMPC_PREC_TC && !enc->tc) return KMPC_ERR_UNSUPPORTED;
  const int nzo = kmpc_encoder_out_dim(enc, lift_mode);
  const int n = enc->p.dims[0];
  if (nzo > KMPC_MAX_NZ) return KMPC_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  WsLock ws(enc, st);
  const int64_t chunk = kGramChunk < M ? kGramChunk : (M > 0 ? M : 1);
  if (ws.acquire((size_t)2 * chunk * nzo) != cudaSuccess) return KMPC_ERR_ALLOC;
  double* px = enc->d_ws;
  double* py = enc->d_ws + (size_t)chunk * nzo;
  for (int64_t m0 = 0; m0 < M; m0 += chunk) {
    const int64_t mc = (M - m0 < chunk) ? (M - m0) : chunk;
    int rc = encode_any(enc, x + m0 * n, px, mc, lift_mode, precision, st);
    if (rc != KMPC_OK) return rc;
    rc = encode_any(enc, y + m0 * n, py, mc, lift_mode, precision, st);
    if (rc != KMPC_OK) return rc;
    rc = kmpc_gram_accumulate(px, py, u + m0, x + m0 * n, mc, nzo, n, pack, stream);
    if (rc != KMPC_OK) return rc;
  }
  return KMPC_OK;
}

int kmpc_gram_from_snapshots(const kmpc_encoder* enc, int lift_mode, const double* x, const double* y,
                             const double* u, int64_t M, double* pack, void* stream) {
  return kmpc_gram_from_snapshots_ex(enc, lift_mode, KMPC_PREC_FP64, x, y, u, M, pack, stream);
}

// Trajectory-aware fused lift + Gram: the M = n_traj * n_step snapshots are trajectory-major and
// CONSECUTIVE (y of snapshot j is x of snapshot j + 1 of the same trajectory, as data_generate.py
// and kmpc_generate_snapshots produce them), so lift(y_j) == lift(x_{j+1}) bit for bit and every
// state needs ONE encode: n_step + 1 per trajectory instead of 2 n_step.
int kmpc_gram_from_trajectories_ex(const kmpc_encoder* enc_c, int lift_mode, int precision, const double* x,
                                   const double* y, const double* u, int64_t n_traj, int n_step,
                                   double* pack, void* stream) {
  if (!enc_c || !x || !y || !u || !pack || n_traj < 0 || n_step < 1) return KMPC_ERR_ARG;
  if (precision != KMPC_PREC_FP64 && precision != KMPC_PREC_TC) return KMPC_ERR_ARG;
  kmpc_encoder* enc = const_cast<kmpc_encoder*>(enc_c);
  if (precision == KMPC_PREC_TC && !enc->tc) return KMPC_ERR_UNSUPPORTED;
  const int nzo = kmpc_encoder_out_dim(enc, lift_mode);
  const int n = enc->p.dims[0];
  if (nzo > KMPC_MAX_NZ) return KMPC_ERR_UNSUPPORTED;
  if (n_step + 1 > kGramChunk) return KMPC_ERR_UNSUPPORTED;
  int64_t tc = kGramChunk / (n_step + 1);                          // trajectories per chunk
  if (tc > n_traj) tc = n_traj > 0 ? n_traj : 1;
  const int64_t chunk_rows = tc * (n_step + 1);
  cudaStream_t st = as_stream(stream);
  WsLock ws(enc, st);
  if (ws.acquire((size_t)chunk_rows * (nzo + n)) != cudaSuccess) return KMPC_ERR_ALLOC;
  double* pz = enc->d_ws;                                  // lifted rows (chunk_rows x nzo)
  double* pin = enc->d_ws + (size_t)chunk_rows * nzo;      // gathered states (chunk_rows x n)
  for (int64_t t0 = 0; t0 < n_traj; t0 += tc) {
    const int64_t nt = (n_traj - t0 < tc) ? (n_traj - t0) : tc;
    const int64_t rows = nt * (n_step + 1), m0 = t0 * n_step;
    traj_rows_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(x + m0 * n, y + m0 * n, n, nt, n_step, pin);
    KMPC_AFTER_LAUNCH();
    int rc = encode_any(enc, pin, pz, rows, lift_mode, precision, st);
    if (rc != KMPC_OK) return rc;
    rc = gram_accumulate_impl(pz, nullptr, u + m0, x + m0 * n, nt * n_step, nzo, n, pack, n_step, stream);
    if (rc != KMPC_OK) return rc;
  }
  return KMPC_OK;
}

int kmpc_gram_from_trajectories(const kmpc_encoder* enc, int lift_mode, const double* x, const double* y,
                                const double* u, int64_t n_traj, int n_step, double* pack, void* stream) {
  return kmpc_gram_from_trajectories_ex(enc, lift_mode, KMPC_PREC_FP64, x, y, u, n_traj, n_step, pack, stream);
}

}  // extern "C"
