// encoder.cuh -- theta_E encoder pieces shared by lift.cu (stand-alone lift kernels) and fused.cu
// (the persistent closed-loop kernel): packed weight layout, TMA bulk-copy / mbarrier helpers and
// the per-tile layer chain on the fp64 tensor path (mma.sync m8n8k4 f64).
//
// Reference: duffing.py:21-29 (`nn.Sequential(Linear(2,100), ReLU, Linear(100,100), ReLU,
// Linear(100,100), ReLU, Linear(100,8))`), Encoder_Tank.m:3-5 (3 layers, nz = 10).
#pragma once
#include <mutex>
#include <vector>

#include "common.cuh"

namespace kmpc {

constexpr int kTileS = 32;         // scenarios (rows) per CTA tile
constexpr int kMmaWarps = 8;
constexpr int kMmaThreads = kMmaWarps * 32;
constexpr int kActStride = 36;     // doubles per activation row k: 32 scenarios + 4 pad (== 4 mod 16)

struct EncParams {
  int n_layers;
  int dims[KMPC_MAX_LAYERS + 1];
  int pad[KMPC_MAX_LAYERS + 1];
  const double* wt[KMPC_MAX_LAYERS];
  const double* b[KMPC_MAX_LAYERS];
  const double* z0;
  const double* packed;            // [W1t | b1 | W2t | b2 | ...] (same layout as the smem copy)
  int woff[KMPC_MAX_LAYERS];       // offset (doubles) of layer l's W block inside `packed`
  int wlen[KMPC_MAX_LAYERS];       // doubles of layer l's W + bias block
  int total_w;                     // doubles in `packed`
  int actw;                        // activation buffer width (max padded layer width)
  int wstride[KMPC_MAX_LAYERS];    // row stride (doubles) of layer l's W^T block, == 4 (mod 16)
  int inpad[KMPC_MAX_LAYERS];      // rows of layer l's W^T block (in rounded up to 4, zero rows)
};

#ifdef __CUDACC__
// ---- TMA bulk-copy / mbarrier helpers (PTX; SASS: UBLKCP, SYNCS) --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

// ReLU that propagates NaN like torch.nn.ReLU / numpy.maximum (fmax would return 0 for NaN)
__device__ __forceinline__ double relu_nan(double v) { return v < 0.0 ? 0.0 : v; }

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Thread 0 of the CTA: initialise one mbarrier per layer and start the TMA bulk copies of the
// packed weight set into `wsm`.  Callers __syncthreads() between the init and the issue.
__device__ __forceinline__ void encoder_weights_init_barriers(const EncParams& p, uint64_t* bars) {
  for (int l = 0; l < p.n_layers; ++l) mbar_init(&bars[l], 1);
  mbar_fence_init();
}
__device__ __forceinline__ void encoder_weights_issue(const EncParams& p, double* wsm, uint64_t* bars) {
  for (int l = 0; l < p.n_layers; ++l) {
    const uint32_t bytes = (uint32_t)p.wlen[l] * 8u;
    mbar_expect_tx(&bars[l], bytes);
    for (uint32_t o = 0; o < bytes; o += 32768u) {
      const uint32_t sz = (bytes - o < 32768u) ? (bytes - o) : 32768u;
      bulk_copy_g2s(reinterpret_cast<char*>(wsm + p.woff[l]) + o,
                    reinterpret_cast<const char*>(p.packed + p.woff[l]) + o, sz, &bars[l]);
    }
  }
}

// every thread that is going to read the weights: wait until all layers have landed (once per kernel)
__device__ __forceinline__ void encoder_weights_wait_all(const EncParams& p, uint64_t* bars) {
  for (int l = 0; l < p.n_layers; ++l) mbar_wait(&bars[l], 0);
}

// k-loop of one layer for a warp that owns NT n-tiles (two m-tiles each): 2 NT DMMAs per k-step
template <int NT>
__device__ __forceinline__ void encoder_kloop(const double* __restrict__ ap, const double* __restrict__ bp,
                                              int kin, int ws, double (&c)[2][4][2]) {
#pragma unroll 2
  for (int k0 = 0; k0 < kin; k0 += 4) {
    const double a0 = ap[k0 * kActStride];
    const double a1 = ap[k0 * kActStride + 8];
    double b[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) b[j] = bp[k0 * ws + 32 * j];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      dmma_m8n8k4(c[0][j][0], c[0][j][1], a0, b[j]);
      dmma_m8n8k4(c[1][j][0], c[1][j][1], a1, b[j]);
    }
  }
}

// All layers of one 32-row tile.  The layer-0 input must already be in `in0` (k-major,
// in0[k * kActStride + row], rows [n, inpad[0]) zero; may be `act` itself) and visible
// (__syncthreads() by the caller).
// Warp w owns rows 16*(w&1)..+15 (two m-tiles) and the n-tiles {w>>1, (w>>1)+4, ...} of every
// layer; activations are rewritten in place between two barriers.  Final outputs are handed to
// store(row_in_tile, col, value) (before the subtraction of theta(0), which the caller applies).
// Ends with a __syncthreads(): `act` is free and the stores of all warps are done.
template <typename Store>
__device__ __forceinline__ void encoder_layers(const EncParams& p, const double* in0, double* act,
                                               const double* wsm, uint64_t* bars, Store store) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;   // mma fragment coordinates
  const int mrow = 16 * (warp & 1);            // first row of this warp's two m-tiles
  const int ng = warp >> 1;                    // n-tile group
  for (int l = 0; l < p.n_layers; ++l) {
    const int kin = p.inpad[l], out = p.dims[l + 1], ws = p.wstride[l];
    const int nt = (out + 7) >> 3;
    const bool last = (l == p.n_layers - 1);
    mbar_wait(&bars[l], 0);  // layer l's weights have landed (returns at once after the first tile)
    const double* wt = wsm + p.woff[l];
    const double* bias = wt + kin * ws;
    double c[2][4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n0 = (ng + 4 * j) * 8;
      const double b0 = (ng + 4 * j < nt) ? bias[n0 + 2 * tig] : 0.0;
      const double b1 = (ng + 4 * j < nt) ? bias[n0 + 2 * tig + 1] : 0.0;
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        c[m][j][0] = b0;
        c[m][j][1] = b1;
      }
    }
    const double* ap = (l == 0 ? in0 : act) + tig * kActStride + mrow + gid;
    const double* bp = wt + tig * ws + ng * 8 + gid;
    // number of n-tiles this warp owns in this layer (warp-uniform): the k-loop is instantiated per
    // count so that no DMMA is predicated (a predicated DMMA costs ~27 instead of 16 cycles)
    const int my_nt = (nt > ng) ? ((nt - ng + 3) >> 2) : 0;
    switch (my_nt) {
      case 4: encoder_kloop<4>(ap, bp, kin, ws, c); break;
      case 3: encoder_kloop<3>(ap, bp, kin, ws, c); break;
      case 2: encoder_kloop<2>(ap, bp, kin, ws, c); break;
      case 1: encoder_kloop<1>(ap, bp, kin, ws, c); break;
      default: break;
    }
    __syncthreads();  // every warp has finished reading the activations of this layer
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (ng + 4 * j < nt) {
        const int col = (ng + 4 * j) * 8 + 2 * tig;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const int r = mrow + 8 * m + gid;
          if (!last) {
            act[col * kActStride + r] = relu_nan(c[m][j][0]);
            act[(col + 1) * kActStride + r] = relu_nan(c[m][j][1]);
          } else {
            if (col < out) store(r, col, c[m][j][0]);
            if (col + 1 < out) store(r, col + 1, c[m][j][1]);
          }
        }
      }
    }
    __syncthreads();
  }
}
// ---- lift-unit variant (fused.cu): a group of W warps lifts ITS unit of 8 scenarios (one m-tile),
// synchronising only inside the group through a named barrier, so the units of a CTA drift freely.
// Warp lw owns the n-tiles lw, lw + W, lw + 2W, ... of every hidden layer; activations ping-pong
// between two k-major buffers of pitch 8 (ONE group barrier per layer); the k-loop is software
// pipelined (the fragments of k-step i + 1 are in flight while the DMMAs of k-step i issue, no DMMA is
// predicated); the last layer (one n-tile) is split over the warps along K, two accumulator chains
// per warp, and summed (+ bias) by the first 64 threads of the group.
constexpr int kUnitRows = 8;       // scenarios per lift unit = rows of one m-tile
constexpr int kActPitch = 8;       // doubles per activation row k
// Activation element (k, row) lives at k * 8 + (row ^ 4 * ((k >> 1) & 1)).  64-bit shared accesses
// are served per half-warp (16 lanes x 8 B = all 32 banks): the A fragment of an m8n8k4 DMMA reads
// rows 0..3 (or 4..7) of four consecutive k in a half-warp, and the XOR puts the k = 2, 3 rows in the
// other half of the 16-double window -- conflict free with no padding (measured 2-way without it).
__host__ __device__ __forceinline__ int act_index(int k, int row) {
  return k * kActPitch + (row ^ (((k >> 1) & 1) << 2));
}

template <int THREADS>
__device__ __forceinline__ void group_barrier(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(THREADS) : "memory");
}

// NT n-tiles (stride 8 W columns) x `ksteps` k-steps; ap/bp point at this lane's first A / B
// fragment element.  Two k-steps per iteration, double-buffered fragments; an odd k-step is peeled.
template <int NT, int W, int MT>
__device__ __forceinline__ void lift_kloop(const double* __restrict__ ap, const double* __restrict__ bp,
                                           int ksteps, int ws, double (&c)[MT][2]) {
  double a0 = ap[0], b0[NT], a1, b1[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) b0[j] = bp[8 * W * j];
  const int pairs = ksteps >> 1, last = ksteps - 1;
  const int astep = 4 * kActPitch, bstep = 4 * ws;
#pragma unroll 1
  for (int i = 0; i < pairs; ++i) {
    const double* ap1 = ap + (2 * i + 1) * astep;
    const double* bp1 = bp + (2 * i + 1) * bstep;
    a1 = ap1[0];
#pragma unroll
    for (int j = 0; j < NT; ++j) b1[j] = bp1[8 * W * j];
#pragma unroll
    for (int j = 0; j < NT; ++j) dmma_m8n8k4(c[j][0], c[j][1], a0, b0[j]);
    const int k2 = min(2 * i + 2, last);   // the very last prefetch re-loads a valid k-step
    const double* ap2 = ap + k2 * astep;
    const double* bp2 = bp + k2 * bstep;
    a0 = ap2[0];
#pragma unroll
    for (int j = 0; j < NT; ++j) b0[j] = bp2[8 * W * j];
#pragma unroll
    for (int j = 0; j < NT; ++j) dmma_m8n8k4(c[j][0], c[j][1], a1, b1[j]);
  }
  if (ksteps & 1) {
#pragma unroll
    for (int j = 0; j < NT; ++j) dmma_m8n8k4(c[j][0], c[j][1], a0, b0[j]);
  }
}

// One unit through all layers.  in0: layer-0 input, k-major in0[act_index(k, row)] (rows [n, inpad[0]) zero);
// bufA / bufB: ping-pong activation buffers (p.actw * kActPitch >= 64 W doubles each; the split-K
// partial sums of the last layer go to the one that layer does not read; in0 and y may live inside
// bufB, outside the first 64 W doubles); y[row * ypitch + col] (before the subtraction of theta(0)).
// bufB == bufA selects IN-PLACE activations (half the shared memory, so twice the units per CTA in the
// stand-alone encoder kernel): one more group barrier per layer between the last read and the first
// write; in0 may then live in the pad rows of bufA (k-rows >= the widest layer) and y anywhere in
// bufA beyond the first 64 W doubles.
// The caller has waited for the weights (encoder_weights_wait_all) and synchronises the group before
// the call (in0 visible, buffers free); ends with a group barrier (y visible, buffers free).
template <int W>
__device__ __forceinline__ void lift_unit(const EncParams& p, const double* in0, double* bufA, double* bufB,
                                          double* y, int ypitch, const double* wsm, int lw, int lane,
                                          int bar_id) {
  constexpr int MT = (KMPC_MAX_WIDTH / 8 + W - 1) / W;   // n-tiles per warp at the widest layer
  const int gid = lane >> 2, tig = lane & 3;   // mma fragment coordinates
  const double* src = in0;
  double* dst = bufA;
  const bool inplace = (bufA == bufB);   // warp-uniform
  const int nl = p.n_layers;
  for (int l = 0; l + 1 < nl; ++l) {
    const int ksteps = p.inpad[l] >> 2, out = p.dims[l + 1], ws = p.wstride[l];
    const int nt = (out + 7) >> 3;
    const double* wt = wsm + p.woff[l];
    const double* bias = wt + p.inpad[l] * ws;
    double c[MT][2];
#pragma unroll
    for (int j = 0; j < MT; ++j) {
      const int tile = lw + W * j;
      const double2 bb = (tile < nt) ? *reinterpret_cast<const double2*>(bias + tile * 8 + 2 * tig)
                                     : make_double2(0.0, 0.0);
      c[j][0] = bb.x;
      c[j][1] = bb.y;
    }
    const double* ap = src + act_index(tig, gid);   // k = 4 ks + tig: the swizzle depends on tig only
    const double* bp = wt + tig * ws + lw * 8 + gid;
    const int my_nt = (nt > lw) ? ((nt - lw + W - 1) / W) : 0;   // warp-uniform
    switch (my_nt) {
      case 8: if (MT >= 8) lift_kloop<(MT >= 8 ? 8 : 1), W, MT>(ap, bp, ksteps, ws, c); break;
      case 7: if (MT >= 7) lift_kloop<(MT >= 7 ? 7 : 1), W, MT>(ap, bp, ksteps, ws, c); break;
      case 6: if (MT >= 6) lift_kloop<(MT >= 6 ? 6 : 1), W, MT>(ap, bp, ksteps, ws, c); break;
      case 5: if (MT >= 5) lift_kloop<(MT >= 5 ? 5 : 1), W, MT>(ap, bp, ksteps, ws, c); break;
      case 4: lift_kloop<4, W, MT>(ap, bp, ksteps, ws, c); break;
      case 3: lift_kloop<3, W, MT>(ap, bp, ksteps, ws, c); break;
      case 2: lift_kloop<2, W, MT>(ap, bp, ksteps, ws, c); break;
      case 1: lift_kloop<1, W, MT>(ap, bp, ksteps, ws, c); break;
      default: break;
    }
    if (inplace) group_barrier<W * 32>(bar_id);   // every warp is done reading what is overwritten now
#pragma unroll
    for (int j = 0; j < MT; ++j) {
      const int tile = lw + W * j;
      if (tile < nt) {
        const int col = tile * 8 + 2 * tig;   // col and col + 1 share (k >> 1) & 1 = tig & 1
        dst[act_index(col, gid)] = relu_nan(c[j][0]);
        dst[act_index(col + 1, gid)] = relu_nan(c[j][1]);
      }
    }
    group_barrier<W * 32>(bar_id);   // layer l + 1 reads what every warp has just written
    src = dst;
    dst = (dst == bufA) ? bufB : bufA;
  }
  // last layer.  One n-tile (out <= 8): K split over the warps, two accumulator chains per warp,
  // partial sums through `dst`.  More n-tiles (Tank: out = 10): tiles over the warps like a hidden
  // layer, no ReLU.  Output y[row * ypitch + col].
  {
    const int l = nl - 1;
    const int ksteps = p.inpad[l] >> 2, out = p.dims[l + 1], ws = p.wstride[l];
    const int nt = (out + 7) >> 3;
    const double* wt = wsm + p.woff[l];
    const double* bias = wt + p.inpad[l] * ws;
    const double* ap = src + act_index(tig, gid);
    if (nt == 1) {
      const int per = (ksteps + W - 1) / W;
      const int kb = min(lw * per, ksteps), ke = min(kb + per, ksteps);
      const double* bp = wt + tig * ws + gid;
      double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll 1
      for (int ks = kb; ks + 1 < ke; ks += 2) {
        const double a0 = ap[ks * (4 * kActPitch)], b0 = bp[ks * 4 * ws];
        const double a1 = ap[(ks + 1) * (4 * kActPitch)], b1 = bp[(ks + 1) * 4 * ws];
        dmma_m8n8k4(c0, c1, a0, b0);
        dmma_m8n8k4(d0, d1, a1, b1);
      }
      if ((ke - kb) & 1) dmma_m8n8k4(c0, c1, ap[(ke - 1) * (4 * kActPitch)], bp[(ke - 1) * 4 * ws]);
      double* part = dst;
      if (inplace) group_barrier<W * 32>(bar_id);
      *reinterpret_cast<double2*>(part + lw * 64 + gid * 8 + 2 * tig) = make_double2(c0 + d0, c1 + d1);
      group_barrier<W * 32>(bar_id);
      const int t = lw * 32 + lane;
      if (t < 64) {
        const int row = t >> 3, col = t & 7;
        double s = part[t];
#pragma unroll
        for (int w = 1; w < W; ++w) s += part[w * 64 + t];
        if (col < out) y[row * ypitch + col] = s + bias[col];
      }
    } else {
      double c[MT][2];
#pragma unroll
      for (int j = 0; j < MT; ++j) {
        const int tile = lw + W * j;
        const double2 bb = (tile < nt) ? *reinterpret_cast<const double2*>(bias + tile * 8 + 2 * tig)
                                       : make_double2(0.0, 0.0);
        c[j][0] = bb.x;
        c[j][1] = bb.y;
      }
      const double* bp = wt + tig * ws + lw * 8 + gid;
      const int my_nt = (nt > lw) ? ((nt - lw + W - 1) / W) : 0;
      switch (my_nt) {   // a last layer wider than 2 W tiles is not a lift (fused_eligible / units_eligible)
        case 2: lift_kloop<2, W, MT>(ap, bp, ksteps, ws, c); break;
        case 1: lift_kloop<1, W, MT>(ap, bp, ksteps, ws, c); break;
        default: break;
      }
      if (inplace) group_barrier<W * 32>(bar_id);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int tile = lw + W * j;
        if (tile < nt) {
          const int col = tile * 8 + 2 * tig;
          if (col < out) y[gid * ypitch + col] = c[j][0];
          if (col + 1 < out) y[gid * ypitch + col + 1] = c[j][1];
        }
      }
    }
    group_barrier<W * 32>(bar_id);
  }
}
#endif  // __CUDACC__

}  // namespace kmpc

namespace kmpc {
struct TcState;   // tc_lift.cu: split-precision weight image + tensor map of the tcgen05 lift
TcState* tc_state_create(const double* const* W, const double* const* b, const int* dims, int n_layers,
                         const double* d_z0, cudaStream_t st);
void tc_state_destroy(TcState* t);
}  // namespace kmpc

// the opaque handle of include/kmpc.h
struct kmpc_encoder {
  kmpc::EncParams p;
  int smem_bytes = 0;      // dynamic smem of encoder_mma_kernel (0: does not fit, use the fallback)
  int num_sms = 148;
  int max_smem_optin = 0;  // cudaDevAttrMaxSharedMemoryPerBlockOptin
  std::vector<double*> owned;
  double* d_z0 = nullptr;
  // L2-resident lift workspace of the fused lift + Gram entry points (allocated on first use).  One
  // workspace per handle: calls on the same handle are serialised -- on the host by ws_mu (held while a
  // call enqueues its work) and on the device by ws_done (recorded when a call's last kernel has been
  // enqueued; the next call's stream waits on it), so two streams / threads never share its contents.
  double* d_ws = nullptr;
  size_t ws_doubles = 0;
  std::mutex ws_mu;
  cudaEvent_t ws_done = nullptr;
  kmpc::TcState* tc = nullptr;   // null: the net does not fit the tcgen05 kernel (KMPC_PREC_TC unsupported)
};
