// tc_lift.cu -- theta_E on the Blackwell tensor path: tcgen05.mma + TMEM + TMA, split precision.
//
// Reference: duffing.py:152-164 (20 000 single-sample `net.Encoder` calls ahead of the EDMD
// regression, duffing.py:167-177) -- the dense contraction of the path.  This is the EDMD-side lift
// (KMPC_PREC_TC); the closed loop keeps the fp64 tensor path (encoder.cuh), because its RLS restart
// transient amplifies 1e-16 to 1e-6 and would amplify the 1e-7 of this kernel to O(1).
//
// Arithmetic.  tcgen05 has no f64 kind.  Every fp32 operand v is split into three bf16 pieces
// v = p0 + p1 + p2 (8 + 8 + 8 significand bits, exact), and a product a*b is evaluated as the six
// piece products of order <= 2:  a2 b0 + a1 b1 + a0 b2 + a1 b0 + a0 b1 + a0 b0  (the dropped ones are
// below 2^-24 |a b|), accumulated in fp32 in TMEM.  The small products are issued FIRST so that only
// the last K/16 accumulations (a0 b0) round at the full magnitude of the sum.  Measured against
// the fp64 path: see tests/test_tc_lift.py (lifted states ~1e-7 relative).
//
// Mapping.  A tile is 128 rows (snapshots) = the 128 TMEM lanes = UMMA M.  Thread t of an epilogue
// warpgroup owns row t: layer 0 (K = 2) runs on CUDA cores in fp64, its ReLU output is split and
// written straight into TMEM as the A operand of the next layer (tcgen05.st; A never touches shared
// memory); hidden layers are D[128 x NP] += A[128 x 16] . B[NP x 16]^T with B = the bf16 weight pieces,
// resident in shared memory for the whole kernel (brought in once per CTA by TMA tensor copies,
// cp.async.bulk.tensor.2d, 128-byte swizzle = the canonical K-major UMMA layout); the epilogue reads
// the accumulator with tcgen05.ld, adds the bias, applies ReLU, splits and stores the next A.
// Warp roles (288 threads): warpgroups 0 / 1 own TMEM slots 0 / 1 (two tiles in flight: while the
// tensor pipe multiplies one tile the other tile's epilogue runs), warp 8 loads the weights and issues
// every tcgen05.mma.  TMEM columns: [0,128) accumulator (shared by both slots: an epilogue drains it
// into registers first and releases it before doing its math), [128,320) A pieces of slot 0,
// [320,512) of slot 1.  All hand-offs are mbarriers (tcgen05.commit for the MMA side).
#include <cuda.h>

#include <string.h>

#include "encoder.cuh"

namespace kmpc {

constexpr int kTcRows = 128;        // rows per tile = UMMA M = TMEM lanes
constexpr int kTcNP = 112;          // padded hidden width: UMMA N of a hidden layer, K of the next
constexpr int kTcKS = kTcNP / 16;   // k-steps (UMMA K = 16 for 16-bit operands)
constexpr int kTcNLast = 16;        // UMMA N of the output layer (out <= 16)
constexpr int kTcKPad = 128;        // K extent of the weight images: two 64-element (128 B) swizzle rows
constexpr int kTcThreads = 288;
constexpr int kTcMaxGemm = 4;       // hidden GEMM layers
constexpr int kTcBlkBytes = kTcNP * 128;            // one K-block of one hidden piece: NP rows x 128 B
constexpr int kTcPieceBytes = 2 * kTcBlkBytes;      // one hidden piece
constexpr int kTcLastBlkBytes = kTcNLast * 128;
constexpr int kTcLastPieceBytes = 2 * kTcLastBlkBytes;
constexpr int kTcAccCol = 0, kTcACol0 = 128, kTcAColStride = 192, kTcCols = 512;

struct TcDev {   // kernel argument (by value)
  int n_in, d1, n_gemm, out;
  const double* w1;        // [NP][6] fp64: layer-0 rows (w0 w1 w2 w3 bias 0), zero rows beyond d1
  const float* bias_h;     // [n_gemm][NP] fp32, zero padded
  const double* bias_o;    // [16] fp64, zero padded
  const double* z0;        // theta(0) from the fp64 path (OFFSET / STACK modes)
  uint32_t w_bytes;        // bytes of the shared-memory weight image
};

struct TcSmem {
  int w, w1, bias_h, bias_o, bars, tmem, total;
};
constexpr int kTcW1Row = 6;   // layer-0 table row: w[0..3] (zero padded), bias, pad -- 48 B, three 16-byte loads
__host__ __device__ inline TcSmem tc_smem_layout(int n_gemm) {
  TcSmem L;
  L.w = 0;
  int o = n_gemm * 3 * kTcPieceBytes + 3 * kTcLastPieceBytes;
  L.w1 = o;
  o += kTcNP * kTcW1Row * 8;
  L.bias_o = o;
  o += 16 * 8;
  L.bias_h = o;
  o += (n_gemm > 0 ? n_gemm : 1) * kTcNP * 4;
  o = (o + 7) & ~7;
  L.bars = o;
  o += 8 * 8;
  L.tmem = o;
  o += 16;
  L.total = o + 1024;   // slack for the 1024-byte alignment of the weight image
  return L;
}

#ifdef __CUDACC__
// ---- PTX wrappers (SASS: UTCHMMA, LDTM/STTM, UTMALDG, UTCBAR) ---------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]^T, kind::f16 (bf16 operands, fp32 accumulate), one CTA
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_desc_lo, uint32_t b_desc_hi,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 bd;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 bd, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_desc_lo), "r"(b_desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (SASS ELECT): the issuer of the warp's tcgen05.mma / TMA instructions
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// Shared-memory matrix descriptor, K-major, 128-byte swizzle: 8-row groups 1024 B apart (SBO), LBO
// unused, descriptor version 1 (sm_100).  Low word: start address >> 4 | LBO; high word: constant.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
// bf16 x bf16 -> f32, A and B K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// (lo, hi) fp32 pair -> three packed bf16x2 words: word p holds piece p of lo (bits 0..15, even k)
// and of hi (bits 16..31, odd k).  The residuals are exact in fp32.
__device__ __forceinline__ void split3(float lo, float hi, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(hi), "f"(lo));
#ifdef KMPC_TC_DIAG_NOSPLIT   // timing diagnostic only (results are bf16-accurate): no residual pieces
  p1 = p2 = 0u;
  return;
#endif
  const float lo1 = lo - __uint_as_float(p0 << 16), hi1 = hi - __uint_as_float(p0 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(hi1), "f"(lo1));
  const float lo2 = lo1 - __uint_as_float(p1 << 16), hi2 = hi1 - __uint_as_float(p1 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(hi2), "f"(lo2));
}
// 16 activations (one k-step) of this thread's row -> the three A pieces in TMEM
__device__ __forceinline__ void store_a_chunk(uint32_t a_lane_base, int ch, const float (&v)[16]) {
  uint32_t w0[8], w1[8], w2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split3(v[2 * i], v[2 * i + 1], w0[i], w1[i], w2[i]);
  tmem_st8(a_lane_base + 0 * (kTcNP / 2) + ch * 8, w0);
  tmem_st8(a_lane_base + 1 * (kTcNP / 2) + ch * 8, w1);
  tmem_st8(a_lane_base + 2 * (kTcNP / 2) + ch * 8, w2);
}

template <int NIN>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_encoder_kernel(const __grid_constant__ CUtensorMap wmap, const TcDev p, const double* __restrict__ x,
                  double* __restrict__ z, int64_t S, int lift_mode, int out_dim, int64_t n_pairs) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned (128 B swizzle atoms)
  const TcSmem L = tc_smem_layout(p.n_gemm);
  double* w1s = reinterpret_cast<double*>(smem + L.w1);
  double* bos = reinterpret_cast<double*>(smem + L.bias_o);
  float* bhs = reinterpret_cast<float*>(smem + L.bias_h);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem);
  uint64_t* wbar = &bars[0];
  uint64_t* a_ready = &bars[1];    // [2]  epilogue -> MMA: the A pieces of the slot are in TMEM
  uint64_t* acc_full = &bars[3];   // [2]  MMA -> epilogue: the accumulator holds the slot's layer
  uint64_t* acc_free = &bars[5];   //      epilogue -> MMA: the accumulator has been drained
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(wbar, 1);
    mbar_init(&a_ready[0], 128);
    mbar_init(&a_ready[1], 128);
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    mbar_init(acc_free, 128);
    mbar_fence_init();
  }
  for (int e = tid; e < kTcNP * kTcW1Row; e += kTcThreads) w1s[e] = p.w1[e];
  for (int e = tid; e < 16; e += kTcThreads) bos[e] = p.bias_o[e];
  for (int e = tid; e < p.n_gemm * kTcNP; e += kTcThreads) bhs[e] = p.bias_h[e];
  if (warp == 8) {   // TMEM: all 512 columns (one CTA per SM: the weight image fills shared memory)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTcCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // pairs of tiles handled by this CTA: the same count in every role
  const int64_t my_iters = (n_pairs > (int64_t)blockIdx.x) ? (n_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int n_batches = p.n_gemm + 1;   // tensor batches per tile: hidden GEMMs + the output layer

  if (warp == 8) {
    // ================= TMA producer + MMA issuer =================
    if (lane == 0) {
      mbar_expect_tx(wbar, p.w_bytes);
      // hidden pieces: rows of the global image are [gemm][piece][NP]; a box is 64 k x 16 rows = 2 KB
      int row = 0;
      for (int m = 0; m < p.n_gemm * 3; ++m)
        for (int r = 0; r < kTcNP; r += 16, row += 16)
          for (int kb = 0; kb < 2; ++kb)
            tma_load_2d(smem + L.w + m * kTcPieceBytes + kb * kTcBlkBytes + (r >> 3) * 1024, &wmap, kb * 64, row, wbar);
      for (int m = 0; m < 3; ++m, row += 16)
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d(smem + L.w + p.n_gemm * 3 * kTcPieceBytes + m * kTcLastPieceBytes + kb * kTcLastBlkBytes, &wmap,
                      kb * 64, row, wbar);
    }
    mbar_wait_bounded(wbar, 0);
    const uint32_t w_lo = umma_desc_lo(smem_u32(smem + L.w));   // descriptor of the first weight byte
    const uint32_t idesc_h = umma_idesc_bf16(kTcNP), idesc_o = umma_idesc_bf16(kTcNLast);
    const bool leader = elect_one();
    // Batches alternate between the slots: X.L1, Y.L1, X.L2, Y.L2, ... (serving them in ready order with a
    // polling loop was measured 12 % slower, staggering the warpgroups made no difference:
    // profiles/r2/tc_variants.log).  done0/1 = batches issued per slot; the layer is done % n_batches.
    const int64_t total = my_iters * n_batches;
    int64_t done0 = 0, done1 = 0;   // scalars (a two-element array indexed by `slot` would live in local memory)
    uint32_t n_batch = 0;
    int slot = 0;
    while (done0 < total || done1 < total) {
      const int64_t d = slot ? done1 : done0;
      mbar_wait_bounded(&a_ready[slot], (uint32_t)d & 1);
      const int j = (int)(d % n_batches);
      if (slot) ++done1;
      else ++done0;
      if (n_batch > 0) mbar_wait_bounded(acc_free, (n_batch - 1) & 1);
      ++n_batch;
      tc_fence_after();
      if (leader) {
        const uint32_t a_base = tmem_base + kTcACol0 + slot * kTcAColStride;
        const uint32_t acc = tmem_base + kTcAccCol;
        // order-2 products first, the dominant a0.b0 last (see the header); every descriptor is
        // the layer's base plus a compile-time offset (16-byte units)
        if (j < p.n_gemm) {
          const uint32_t wl = w_lo + (uint32_t)j * (3 * kTcPieceBytes >> 4);
#pragma unroll
          for (int t = 0; t < 6; ++t) {
            const int pa = (t == 0) ? 2 : ((t == 1 || t == 3) ? 1 : 0);
            const int pb = (t == 2) ? 2 : ((t == 1 || t == 4) ? 1 : 0);
#pragma unroll
            for (int ks = 0; ks < kTcKS; ++ks)
              umma_ts(acc, a_base + pa * (kTcNP / 2) + ks * 8,
                      wl + ((pb * kTcPieceBytes + (ks >> 2) * kTcBlkBytes + (ks & 3) * 32) >> 4), kDescHiSw128,
                      idesc_h, (t | ks) ? 1u : 0u);
          }
        } else {
          const uint32_t wl = w_lo + (uint32_t)p.n_gemm * (3 * kTcPieceBytes >> 4);
#pragma unroll
          for (int t = 0; t < 6; ++t) {
            const int pa = (t == 0) ? 2 : ((t == 1 || t == 3) ? 1 : 0);
            const int pb = (t == 2) ? 2 : ((t == 1 || t == 4) ? 1 : 0);
#pragma unroll
            for (int ks = 0; ks < kTcKS; ++ks)
              umma_ts(acc, a_base + pa * (kTcNP / 2) + ks * 8,
                      wl + ((pb * kTcLastPieceBytes + (ks >> 2) * kTcLastBlkBytes + (ks & 3) * 32) >> 4),
                      kDescHiSw128, idesc_o, (t | ks) ? 1u : 0u);
          }
        }
        umma_commit(&acc_full[slot]);
      }
      __syncwarp();
      slot ^= 1;   // look at the other slot first next time
    }
  } else {
    // ================= epilogue warpgroups: slot = warpgroup =================
    const int slot = warp >> 2, row_in = tid & 127;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t a_lane_base = lane_base + kTcACol0 + slot * kTcAColStride;
    const uint32_t acc_lane_base = lane_base + kTcAccCol;
    const int n = p.n_in, off = (lift_mode == KMPC_LIFT_STACK) ? n : 0;
    uint32_t full_cnt = 0;
    // this thread's row of a tile: x is prefetched one tile ahead (the load is in flight during the
    // whole previous tile instead of stalling layer 0)
    auto load_x = [&](int64_t it, double (&xv)[4]) {
      const int64_t r = (((int64_t)blockIdx.x + it * gridDim.x) * 2 + slot) * kTcRows + row_in;
#pragma unroll
      for (int k = 0; k < NIN; ++k) xv[k] = (k < n && it < my_iters && r < S) ? __ldg(x + r * n + k) : 0.0;
    };
    double xnext[4] = {0.0, 0.0, 0.0, 0.0};
    load_x(0, xnext);
    for (int64_t it = 0; it < my_iters; ++it) {
      const int64_t tile = ((int64_t)blockIdx.x + it * gridDim.x) * 2 + slot;
      const int64_t row = tile * kTcRows + row_in;
      const bool valid = row < S;
      // ---- layer 0 on CUDA cores (fp64): h = relu(W1 x + b1) -> A pieces ----
      double xin[4] = {xnext[0], xnext[1], xnext[2], xnext[3]};
      load_x(it + 1, xnext);
      if (valid && lift_mode == KMPC_LIFT_STACK) {
#pragma unroll
        for (int k = 0; k < NIN; ++k)
          if (k < n) z[row * out_dim + k] = xin[k];
      }
      // branch-free: the table is zero beyond d1 / n_in, so padded outputs come out as relu(0) = 0
#pragma unroll 1
      for (int ch = 0; ch < kTcKS; ++ch) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const double2* wr = reinterpret_cast<const double2*>(w1s + (ch * 16 + i) * kTcW1Row);
          const double2 w01 = wr[0], bb = wr[2];
#ifdef KMPC_TC_DIAG_L0F32   // timing diagnostic only (results lose 1e-7): layer 0 in fp32
          float f = fmaf((float)w01.x, (float)xin[0], (float)bb.x);
          f = fmaf((float)w01.y, (float)xin[1], f);
#else
          double h = fma(w01.x, xin[0], bb.x);
          h = fma(w01.y, xin[1], h);
          if (NIN > 2) {
            const double2 w23 = wr[1];
            h = fma(w23.x, xin[2], h);
            h = fma(w23.y, xin[3], h);
          }
          const float f = (float)h;
#endif
          v[i] = f < 0.f ? 0.f : f;   // NaN-propagating ReLU (rounding and ReLU commute)
        }
        store_a_chunk(a_lane_base, ch, v);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&a_ready[slot]);
      // ---- hidden GEMM epilogues: drain the accumulator, release it, bias + ReLU + split -> A ----
      for (int j = 0; j < p.n_gemm; ++j) {
        mbar_wait_bounded(&acc_full[slot], full_cnt & 1);
        ++full_cnt;
        tc_fence_after();
        uint32_t acc[kTcKS][16];
#pragma unroll
        for (int ch = 0; ch < kTcKS; ++ch) tmem_ld16(acc_lane_base + ch * 16, acc[ch]);
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(acc_free);
        const float* bj = bhs + j * kTcNP;
#pragma unroll
        for (int ch = 0; ch < kTcKS; ++ch) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float s = __uint_as_float(acc[ch][i]) + bj[ch * 16 + i];
            v[i] = s < 0.f ? 0.f : s;   // NaN-propagating ReLU
          }
          store_a_chunk(a_lane_base, ch, v);
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&a_ready[slot]);
      }
      // ---- output layer ----
      mbar_wait_bounded(&acc_full[slot], full_cnt & 1);
      ++full_cnt;
      tc_fence_after();
      uint32_t o16[16];
      tmem_ld16(acc_lane_base, o16);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(acc_free);
      if (valid) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          if (c < p.out) {
            double v = (double)__uint_as_float(o16[c]) + bos[c];
            if (lift_mode != KMPC_LIFT_RAW) v -= p.z0[c];
            z[row * out_dim + off + c] = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcCols) : "memory");
  }
}
#endif  // __CUDACC__

// ---------------------------------------------------------------------------- host side -------
struct TcState {
  bool ok = false;
  TcDev dev{};
  CUtensorMap wmap;
  void* d_w = nullptr;        // bf16 weight image [rows][128]
  void* d_misc = nullptr;     // w1 | b1 | bias_o | bias_h
  int smem_bytes = 0;
};

static uint16_t bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);   // inf / nan
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf16_to_f(uint16_t h) {
  const uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

void tc_state_destroy(TcState* t) {
  if (!t) return;
  if (t->d_w) cudaFree(t->d_w);
  if (t->d_misc) cudaFree(t->d_misc);
  delete t;
}

// Build the split-precision weight image for an encoder (nn.Linear layout W[l] (out, in)).  Returns
// nullptr when the net does not fit the kernel's shape (hidden width > 112, out > 16, in > 4, fewer
// than 2 layers) or the driver lacks cuTensorMapEncodeTiled; the tensor path is then reported as
// unsupported by the *_ex entry points (it never silently falls back).
TcState* tc_state_create(const double* const* W, const double* const* b, const int* dims, int n_layers,
                         const double* d_z0, cudaStream_t st) {
  if (n_layers < 2 || n_layers - 2 > kTcMaxGemm) return nullptr;
  if (dims[0] > 4 || dims[n_layers] > kTcNLast) return nullptr;
  for (int l = 1; l < n_layers; ++l)
    if (dims[l] > kTcNP) return nullptr;
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return nullptr;
  const int n_gemm = n_layers - 2, d1 = dims[1], n_in = dims[0], out = dims[n_layers];
  const int rows = n_gemm * 3 * kTcNP + 3 * kTcNLast;
  std::vector<uint16_t> img((size_t)rows * kTcKPad, 0);
  auto put = [&](int row0, int nrows, const double* Wl, int o_dim, int i_dim) {
    for (int o = 0; o < o_dim && o < nrows; ++o)
      for (int k = 0; k < i_dim; ++k) {
        const float f = (float)Wl[(size_t)o * i_dim + k];
        const uint16_t p0 = bf16_rn(f);
        const float r1 = f - bf16_to_f(p0);
        const uint16_t p1 = bf16_rn(r1);
        const float r2 = r1 - bf16_to_f(p1);
        const uint16_t p2 = bf16_rn(r2);
        img[(size_t)(row0 + 0 * nrows + o) * kTcKPad + k] = p0;
        img[(size_t)(row0 + 1 * nrows + o) * kTcKPad + k] = p1;
        img[(size_t)(row0 + 2 * nrows + o) * kTcKPad + k] = p2;
      }
  };
  for (int g = 0; g < n_gemm; ++g) put(g * 3 * kTcNP, kTcNP, W[g + 1], dims[g + 2], dims[g + 1]);
  put(n_gemm * 3 * kTcNP, kTcNLast, W[n_layers - 1], out, dims[n_layers - 1]);
  // misc block: layer-0 table [NP][6] f64 | bias_o [16] f64 | bias_h [n_gemm][NP] f32
  const size_t n_w1 = (size_t)kTcNP * kTcW1Row, misc_d = n_w1 + 16;
  std::vector<double> md(misc_d, 0.0);
  std::vector<float> mf((size_t)(n_gemm > 0 ? n_gemm : 1) * kTcNP, 0.f);
  for (int o = 0; o < d1; ++o) {
    for (int k = 0; k < n_in; ++k) md[(size_t)o * kTcW1Row + k] = W[0][(size_t)o * n_in + k];
    md[(size_t)o * kTcW1Row + 4] = b[0][o];
  }
  for (int o = 0; o < out; ++o) md[n_w1 + o] = b[n_layers - 1][o];
  for (int g = 0; g < n_gemm; ++g)
    for (int o = 0; o < dims[g + 2]; ++o) mf[(size_t)g * kTcNP + o] = (float)b[g + 1][o];
  TcState* t = new TcState();
  auto fail = [&]() {
    cudaGetLastError();
    tc_state_destroy(t);
    return (TcState*)nullptr;
  };
  if (cudaMalloc(&t->d_w, img.size() * 2) != cudaSuccess) return fail();
  if (cudaMalloc(&t->d_misc, md.size() * 8 + mf.size() * 4) != cudaSuccess) return fail();
  if (cudaMemcpyAsync(t->d_w, img.data(), img.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(t->d_misc, md.data(), md.size() * 8, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync((char*)t->d_misc + md.size() * 8, mf.data(), mf.size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess)
    return fail();
  const cuuint64_t gdim[2] = {(cuuint64_t)kTcKPad, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)kTcKPad * 2};
  const cuuint32_t box[2] = {64, 16};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&t->wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, t->d_w, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return fail();
  t->dev.n_in = n_in;
  t->dev.d1 = d1;
  t->dev.n_gemm = n_gemm;
  t->dev.out = out;
  const double* dm = reinterpret_cast<const double*>(t->d_misc);
  t->dev.w1 = dm;
  t->dev.bias_o = dm + n_w1;
  t->dev.bias_h = reinterpret_cast<const float*>(dm + misc_d);
  t->dev.z0 = d_z0;
  t->dev.w_bytes = (uint32_t)(n_gemm * 3 * kTcPieceBytes + 3 * kTcLastPieceBytes);
  t->smem_bytes = tc_smem_layout(n_gemm).total;
  int dev = 0, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (t->smem_bytes > max_smem) return fail();
  // the opt-in limit is a per-function attribute shared by every encoder handle of the process: raise it
  // to the device maximum once (a per-handle value would be overwritten by the next, smaller net)
  if (cudaFuncSetAttribute(tc_encoder_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess ||
      cudaFuncSetAttribute(tc_encoder_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess)
    return fail();
  t->ok = true;
  return t;
}

int tc_encode_launch(const kmpc_encoder* enc, const double* x, double* z, int64_t S, int lift_mode, cudaStream_t st) {
  const TcState* t = enc->tc;
  if (!t || !t->ok) return KMPC_ERR_UNSUPPORTED;
  const int64_t tiles = (S + kTcRows - 1) / kTcRows, pairs = (tiles + 1) / 2;
  const unsigned grid = (unsigned)(pairs < enc->num_sms ? pairs : enc->num_sms);
  const int out_dim = kmpc_encoder_out_dim(enc, lift_mode);
  if (t->dev.n_in <= 2)
    tc_encoder_kernel<2><<<grid, kTcThreads, t->smem_bytes, st>>>(t->wmap, t->dev, x, z, S, lift_mode, out_dim, pairs);
  else
    tc_encoder_kernel<4><<<grid, kTcThreads, t->smem_bytes, st>>>(t->wmap, t->dev, x, z, S, lift_mode, out_dim, pairs);
  KMPC_AFTER_LAUNCH();
  return KMPC_OK;
}

}  // namespace kmpc
