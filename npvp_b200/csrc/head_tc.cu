// The two 7x7 reflect-padded convolutions at the ends of the autoencoder on tcgen05: the HEAD (Cin = 32 / 64 -> Cout <= 3,
// ResNetAutoEncoder.py:184-189; first kernel of this file) and the STEM (Cin = 1 / 3 -> Cout = 32 / 64, ResNetAutoEncoder.py:62-66;
// second kernel, same flattened-row scheme with the taps' roles swapped - see its own header below).
//
// ---- head ----
// The mma.sync head (autoencoder_kernels.cu) stages an 8 x 64 output tile, computes, stores: 673 us per 640 Cityscapes frames
// against 125 us of HBM time (r02 ncu: 30 % issue-active, the tile load and the HMMA chain never overlap).  This kernel
// streams whole padded image rows through a shared-memory ring and keeps the same algebra
//     Z[p, (kx,co)] = sum_{ky,ci} X[row(p) + ky, col(p), ci] * W[ky,kx,ci,co]     tcgen05.mma, M = 128 pixels, N = 32, K = 7 x Cin
//     out[y, x, co] = bias[co] + sum_kx Z[p(y, x) + kx, (kx,co)]                  7-term shift-sum in the epilogue
// where p runs over the FLATTENED reflect-padded rows (pitch P = W + 6 rounded up to 8 pixels, so every 128-pixel tile and
// every tap's operand start on a swizzle-atom boundary): the A operand of tap ky is simply the same 128 pixels one padded
// row further down, i.e. an UMMA descriptor whose start address moved by one row pitch.  Each input row is fetched ONCE per
// CTA (one TMA box: Cin x P pixels, x from -3, out-of-range pixels zero-filled, 64 B / 128 B swizzle = one pixel per swizzle
// row) and read by the tensor core seven times from shared memory; a helper warp overwrites the six padding pixels with their
// mirror images (generic-proxy copies + fence.proxy.async) before the row is handed to the MMA warp.
// A tile may straddle the ring's wrap-around, so the first D rows of the ring are loaded a second time behind its end.
// A CTA owns a contiguous range of output rows of the whole batch (balanced to one row); every frame segment in that range
// costs six extra input rows.  Roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator + padding, warps 4-11
// epilogue (tcgen05.ld -> Z ring in shared memory -> shift-sum -> bias / Tanh | Sigmoid -> fp32 NCHW and / or uint8 pixels).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int kHtRing = 16;            // input rows resident (a tile needs at most D + 7 of them, the rest is prefetch depth)
constexpr int kHtMaxDup = 3;
constexpr int kHtTmemBufs = 8;         // 32-column accumulators in flight
constexpr int kHtZRing = 384;          // Z rows kept for the shift-sum: three tiles
constexpr int kHtZStride = 21;         // floats per Z row (7 kx * 3 co), odd: conflict-free row-per-thread writes and shifted reads
constexpr int kHtThreads = 512;        // warp 0 producer, 1 / 3 MMA issuers, 2 padding + TMEM, 4..15 epilogue
constexpr int kHtEpiThreads = 384;     // 12 warps = 4 TMEM lane quadrants x 3 column groups; the output phase is one (position, channel) per thread
constexpr int kHtMaxRows = 1280;       // stream rows per CTA (row table in shared memory)
constexpr int kHtMaxSeg = 24;

struct HeadTcPix { float inv_std[3], inv_mean[3]; };

struct HeadTcParams {
  const h16* w;          // packed mma.sync fragments (pack_head_weights)
  const float* bias;
  float* out;
  uint8_t* out_u8;
  int64_t total_rows;    // frames * H
  int H, W, P, D, R, Cout, NT, act, fp16;   // P padded row pitch (pixels), D rows mirrored behind the ring, R ring rows
  HeadTcPix pix;
  int wait_mode;
};

// Waits: mode 1 = mbarrier.try_wait (the hardware parks the warp), mode 0 = spin on the non-blocking phase test.  A spinning
// producer / padding / issuer warp shares its scheduler with three epilogue warps and competes for their issue slots.
// Bounded either way: a lost arrival traps instead of hanging the GPU.  (npvp_set_option("head_tc_wait", m), A/B switch.)
__device__ __forceinline__ void ht_wait(uint64_t* bar, uint32_t parity, int mode) {
  if (mode) { ptx::mbar_wait(bar, parity); return; }
  const uint32_t a = ptx::smem_u32(bar);
  uint32_t spins = 0;
  while (!ptx::mbar_test(a, parity)) {
    if (++spins > 400000000u) { printf("npvp head_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void ht_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ht_tma_load_3d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_dst), "l"(map), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void ht_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ht_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
// wait::ld that "redefines" the registers of a tcgen05.ld issued a loop iteration ahead (see tmem_ld_wait_regs in tc_ptx.cuh)
__device__ __forceinline__ void ht_tmem_ld_wait8(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :: "memory");
}
// Tanh / Sigmoid from one ex2 and one reciprocal (absolute error < 3e-7: the frames are quantised to 8 bits downstream)
__device__ __forceinline__ float ht_tanh(float x) { return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(x * 2.885390081777927f)), 1.0f); }
__device__ __forceinline__ float ht_sigmoid(float x) { return rcp_approx(1.0f + ex2_approx(x * -1.4426950408889634f)); }
// All K steps of one tap (KSTEPS = 2 for Cin 32, 4 for Cin 64) from ONE asm statement: descriptors are assembled inside it from
// their low words (+2 = 32 bytes per K step) and a constant high word, so the compiler emits no per-operand uniform-register
// shuffles between the MMAs (r02 ncu: with one statement per MMA and a `lane == 0` branch the two issuing threads were busy 100 % of
// the time - ELECT / R2UR.BROADCAST sequences around every tcgen05.mma - and the tensor pipe only 75 %).
template <int KSTEPS>
__device__ __forceinline__ void ht_umma_tap(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
  if (KSTEPS == 2) {
    asm volatile(
        "{\n\t.reg .pred p, pt;\n\t.reg .b64 da, db;\n\t.reg .b32 la, lb;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.b32 pt, %3, %3;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
        "add.u32 la, %1, 2;\n\tadd.u32 lb, %2, 2;\n\tmov.b64 da, {la, %3};\n\tmov.b64 db, {lb, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, pt;\n\t.reg .b64 da, db;\n\t.reg .b32 la, lb;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.b32 pt, %3, %3;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
        "add.u32 la, %1, 2;\n\tadd.u32 lb, %2, 2;\n\tmov.b64 da, {la, %3};\n\tmov.b64 db, {lb, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
        "add.u32 la, %1, 4;\n\tadd.u32 lb, %2, 4;\n\tmov.b64 da, {la, %3};\n\tmov.b64 db, {lb, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t"
        "add.u32 la, %1, 6;\n\tadd.u32 lb, %2, 6;\n\tmov.b64 da, {la, %3};\n\tmov.b64 db, {lb, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pt;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ int ht_reflect(int i, int n) {       // ReflectionPad2d
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}
// K-major operand descriptor, one pixel (Cin 16-bit values) per swizzle row: SWIZZLE_64B (Cin 32) / SWIZZLE_128B (Cin 64),
// 8-row groups PIXB * 8 bytes apart, descriptor version 1
template <int PIXB>
__device__ __forceinline__ uint64_t ht_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((PIXB * 8) >> 4) << 32) | ((uint64_t)1 << 46);
  return d | ((uint64_t)(PIXB == 128 ? 2 : 4) << 61);
}

__host__ __device__ constexpr uint32_t ht_idesc(int M, int N, int fp16) {      // kind::f16: D = f32, A / B f16 or bf16, K-major
  return (1u << 4) | (fp16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct HtSeg { int s0, n, f, y0; };            // stream rows [s0, s0 + n + 6) produce the n output rows y0 .. y0 + n - 1 of frame f

__device__ __forceinline__ void ht_sts(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float ht_lds(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }

template <int CIN>
__global__ void __launch_bounds__(kHtThreads, 1)
conv7x7_head_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const HeadTcParams p) {
  constexpr int PIXB = CIN * 2;                                   // bytes per pixel = one swizzle row
  constexpr int KSTEPS = CIN / 16;
  extern __shared__ __align__(1024) uint8_t ht_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)ht_smem_raw + 1023) & ~(uintptr_t)1023);
  const int P = p.P, D = p.D, R = p.R, H = p.H, W = p.W, Cout = p.Cout;
  const uint32_t row_bytes = (uint32_t)P * PIXB;                  // a multiple of the swizzle atom (P % 8 == 0)
  uint8_t* ring = smem;                                           // [R + D rows][P pixels][PIXB], swizzled
  uint8_t* wsm = ring + (size_t)(R + D) * row_bytes;              // [7 ky][32 n][PIXB], swizzled
  float* zs = reinterpret_cast<float*>(wsm + 7 * 32 * PIXB);      // [kHtZRing][kHtZStride]
  uint64_t* full = reinterpret_cast<uint64_t*>(zs + kHtZRing * kHtZStride + 2);   // TMA landed            [kHtRing]
  uint64_t* ready = full + kHtRing;                               // padding pixels filled [kHtRing]
  uint64_t* freeb = ready + kHtRing;                              // row consumed by MMAs  [kHtRing]
  uint64_t* tfull = freeb + kHtRing;                              // accumulator complete  [kHtTmemBufs]
  uint64_t* tempty = tfull + kHtTmemBufs;                         // accumulator drained   [kHtTmemBufs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kHtTmemBufs);
  HtSeg* segs = reinterpret_cast<HtSeg*>(tmem_slot + 2);
  int* seg_count = reinterpret_cast<int*>(segs + kHtMaxSeg + 1);
  int64_t* rowtab = reinterpret_cast<int64_t*>(seg_count + 2);   // [n_rows] output offset of (frame, channel 0, row y, column 0), -1: no output row

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this CTA's output rows [o_begin, o_end) of the batch and its row stream
  const int64_t o_begin = p.total_rows * blockIdx.x / gridDim.x, o_end = p.total_rows * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    int ns = 0, s = 0;
    for (int64_t o = o_begin; o < o_end && ns < kHtMaxSeg;) {
      const int y = (int)(o % H);
      const int n = (int)min((int64_t)(H - y), o_end - o);
      segs[ns].s0 = s; segs[ns].n = n; segs[ns].f = (int)(o / H); segs[ns].y0 = y;
      s += n + 6; o += n; ++ns;
    }
    segs[ns].s0 = s; segs[ns].n = 0; segs[ns].f = 0; segs[ns].y0 = 0;     // sentinel: total stream rows
    *seg_count = ns;
    for (int i = 0; i < kHtRing; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&ready[i], 1); ptx::mbar_init(&freeb[i], 1); }
    for (int i = 0; i < kHtTmemBufs; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], kHtEpiThreads / 32); }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmap_x);
  }
  {
    // weights: mma.sync fragment packing (pack_head_weights) -> UMMA B operand, rows n = kx * Cout + co (zero beyond NT * 8), K = ci.
    // All loads of a thread are issued before its first store (one exposed memory latency instead of 19).
    constexpr int PER = (7 * 32 * CIN + kHtThreads - 1) / kHtThreads;
    h16 v[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int i = threadIdx.x + u * kHtThreads;
      const int ci = i % CIN, nn = (i / CIN) % 32, ky = i / (CIN * 32);
      const int n = (nn >> 3) * 7 + (nn & 7);                       // accumulator column nn = 8 g + j holds GEMM column n = 7 g + j (j < 7, g < 3)
      v[u] = 0;
      if (ky < 7 && nn < 24 && (nn & 7) < 7 && n < 7 * Cout) {
        const int pass = ci >> 5, half = (ci >> 4) & 1, kk = ci & 15;
        const int tig = (kk & 7) >> 1, e = (kk & 1) + ((kk >> 3) << 1);
        v[u] = __ldg(p.w + ((((size_t)pass * 14 + ky * 2 + half) * p.NT + (n >> 3)) * 32 + (n & 7) * 4 + tig) * 4 + e);
      }
    }
    const uint32_t w_s = ptx::smem_u32(wsm);
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int i = threadIdx.x + u * kHtThreads;
      const int ci = i % CIN, r = i / CIN, c = ci >> 3;
      const int sw = PIXB == 128 ? (r & 7) : ((r >> 1) & 3);
      if (r < 7 * 32) asm volatile("st.shared.u16 [%0], %1;" ::"r"(w_s + (uint32_t)r * PIXB + ((c ^ sw) << 4) + (ci & 7) * 2), "h"(v[u]) : "memory");
    }
  }
  ptx::fence_proxy_async();
  if (warp == 2) { ptx::tmem_alloc(tmem_slot, kHtTmemBufs * 32); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nseg = *seg_count;
  const int n_rows = segs[nseg].s0;                               // stream rows of this CTA
  // flattened positions [0, (n_rows - 6) * P) carry every valid output
  const int n_tiles = n_rows > 6 ? (int)(((int64_t)(n_rows - 6) * P + 127) / 128) : 0;
  const uint32_t ring_s = ptx::smem_u32(ring);
  for (int k = 0; k < nseg; ++k)
    for (int j = threadIdx.x; j < segs[k].n + 6; j += kHtThreads)
      rowtab[segs[k].s0 + j] = j < segs[k].n ? (((int64_t)segs[k].f * Cout) * H + segs[k].y0 + j) * W : -1;
  __syncthreads();

  // Every role below is ONE dependent instruction stream per tile or row, so the loops carry (slot, phase) pairs forward with
  // compares and adds only: a `% R` / `/ R` pair per row cost ~200 clk in the first version (r02 trace).
  if (warp == 0) {
    // ---------------- TMA producer: one box per stream row (two for the rows mirrored behind the ring) ----------------
    if (ptx::elect_one()) {
      int seg = 0, seg_next = segs[1].s0, f = segs[0].f, yrel = segs[0].y0 - 3, slot = 0;
      uint32_t wrapped = 0, par = 0;                              // par: parity of the phase of freeb[slot] that frees it for this use
      for (int s = 0; s < n_rows; ++s) {
        if (s >= seg_next) { ++seg; seg_next = segs[seg + 1].s0; f = segs[seg].f; yrel = segs[seg].y0 - 3; }
        const int yy = ht_reflect(yrel++, H);
        if (wrapped) ht_wait(&freeb[slot], par, p.wait_mode);
        const bool dup = slot < D;
        ptx::mbar_arrive_expect_tx(&full[slot], dup ? 2 * row_bytes : row_bytes);
        ht_tma_load_3d(ring_s + slot * row_bytes, &tmap_x, &full[slot], 0, -3, f * H + yy);
        if (dup) ht_tma_load_3d(ring_s + (R + slot) * row_bytes, &tmap_x, &full[slot], 0, -3, f * H + yy);
        if (++slot == R) { slot = 0; par ^= wrapped; wrapped = 1; }
      }
    }
  } else if (warp == 2) {
    // ---------------- padding pixels: x' = 0..2 <- 6, 5, 4 and x' = W+3..W+5 <- W+1, W, W-1 (x' = x + 3) ----------------
    const int j = lane >> 2, c = lane & 3;                        // lanes 0..23: padding pixel j, chunk group c
    const int xd = j < 3 ? j : W + j, xs = j < 3 ? 6 - j : W + 4 - j;
    const int sd = PIXB == 128 ? (xd & 7) : ((xd >> 1) & 3), ss = PIXB == 128 ? (xs & 7) : ((xs >> 1) & 3);
    int slot = 0;
    uint32_t par = 0;
    for (int s = 0; s < n_rows; ++s) {
      ht_wait(&full[slot], par, p.wait_mode);
      if (j < 6) {
        for (int rep = 0; rep < (slot < D ? 2 : 1); ++rep) {
          const uint32_t row = ring_s + (uint32_t)(rep ? R + slot : slot) * row_bytes;
#pragma unroll
          for (int cc = 0; cc < PIXB / 64; ++cc) {
            const int ch = c * (PIXB / 64) + cc;
            uint32_t a0, a1, a2, a3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(row + (uint32_t)xs * PIXB + ((ch ^ ss) << 4)) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (uint32_t)xd * PIXB + ((ch ^ sd) << 4)), "r"(a0), "r"(a1), "r"(a2), "r"(a3) : "memory");
          }
        }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ht_mbar_arrive(&ready[slot]);
      if (++slot == R) { slot = 0; par ^= 1; }
    }
  } else if (warp == 1 || warp == 3) {
    // ---------------- two MMA issuers, alternate tiles: per tile 7 taps x KSTEPS MMAs (M 128, N 32, K 16) ----------------
    // (tools/ubench/umma_shapes.cu: an MMA of N <= 64 occupies the tensor pipe for 61 clk whatever its N, so a tile is 14 x 61 clk;
    //  one issuer spent another ~1700 clk per tile on barriers and commits - two issuers keep the pipe fed)
    if (ptx::elect_one()) {
      const int me = warp >> 1;                                   // 0 / 1
      const uint32_t idesc = ht_idesc(128, 32, p.fp16);
      const uint32_t hi = (uint32_t)(ht_desc<PIXB>(0) >> 32);
      const uint32_t a_base = (uint32_t)ht_desc<PIXB>(ring_s), b_base = (uint32_t)ht_desc<PIXB>(ptx::smem_u32(wsm));
      const uint32_t row16 = row_bytes >> 4;
      int rows_ready = 0, rdy_slot = 0;
      uint32_t rdy_par = 0;
      int r0 = 0, off = 0, slot0 = 0;                             // first stream row of the tile, pixel offset in it, its ring slot
      if (me) { off = 128; while (off >= P) { off -= P; ++r0; if (++slot0 == R) slot0 = 0; } }
      uint32_t te_par = 0;                                        // parity that frees the accumulator for its next use
      for (int t = me; t < n_tiles; t += 2) {
        int r_last = r0 + 6, o2 = off + 127;
        while (o2 >= P) { o2 -= P; ++r_last; }
        r_last = min(r_last, n_rows - 1);
        while (rows_ready <= r_last) {
          ht_wait(&ready[rdy_slot], rdy_par, p.wait_mode);
          ++rows_ready;
          if (++rdy_slot == R) { rdy_slot = 0; rdy_par ^= 1; }
        }
        const int buf = t % kHtTmemBufs;
        if (t >= kHtTmemBufs) ht_wait(&tempty[buf], te_par, p.wait_mode);
        if (buf >= kHtTmemBufs - 2 && t >= kHtTmemBufs) te_par ^= 1;   // this issuer's last buffer of a round
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 32;
        const uint32_t a_off = a_base + (uint32_t)off * (PIXB / 16);
        int slot = slot0;
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
          const uint32_t a_lo = a_off + (uint32_t)slot * row16, b_lo = b_base + ky * (32 * PIXB / 16);
          if (++slot == R) slot = 0;
          ht_umma_tap<KSTEPS>(d_tmem, a_lo, b_lo, hi, idesc, ky ? 1u : 0u);
        }
        ptx::umma_commit(&tfull[buf]);
        off += 256;
        while (off >= P) { off -= P; ++r0; if (++slot0 == R) slot0 = 0; }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: 12 warps = 4 TMEM lane quadrants x 3 column groups (8 accumulator columns = 7 GEMM columns each) ----------------
    const int ew = warp - 4, quad = warp & 3, grp = ew >> 2;
    const int m = quad * 32 + lane;                               // tile row held by this thread after tcgen05.ld
    const int e = ew * 32 + lane, m2 = e & 127, co = e >> 7;      // output phase: position m2 of the window, channel co
    const float bias = co < Cout ? __ldg(p.bias + co) : 0.f;
    const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + grp * 8;
    const uint32_t zs_s = ptx::smem_u32(zs), rt_s = ptx::smem_u32(rowtab);
    const size_t plane = (size_t)co * H * W;
    // incremental coordinates (no division in the tile loop): the window of tile t holds the flattened positions 128 t + m2 - 6
    int s = 0, xo = m2 - 6 - 128;                                 // stream row / column of this thread's position, one tile behind
    int zb = (m2 - 6 - 128 + 2 * kHtZRing) % kHtZRing;            // its row in the Z ring
    int zw = (m + kHtZRing - 128) % kHtZRing;                     // Z ring row this thread writes
    int fr_row = 0, fr_off = 0, fr_slot = 0;                      // (thread 128) rows released so far, position of the next tile
    uint32_t tf_par = 0;
    uint32_t r[8];
    if (n_tiles > 0) {
      ht_wait(&tfull[0], 0, p.wait_mode);
      ptx::tc_fence_after();
      ht_tmem_ld8(t_addr, r);
    }
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t % kHtTmemBufs;
      ht_tmem_ld_wait8(r);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ht_mbar_arrive(&tempty[buf]);
      if (threadIdx.x == 128) {
        // every MMA of the tiles <= t has completed: the stream rows in front of tile t + 1 are free (plain arrivals: a
        // tcgen05.commit per freed row cost the issuing thread ~400 clk)
        fr_off += 128;
        while (fr_off >= P) { fr_off -= P; if (fr_row < n_rows) ht_mbar_arrive(&freeb[fr_slot]); ++fr_row; if (++fr_slot == R) fr_slot = 0; }
      }
      zw += 128; if (zw >= kHtZRing) zw -= kHtZRing;
      {
        const uint32_t zr = zs_s + (uint32_t)(zw * kHtZStride + grp * 7) * 4;
#pragma unroll
        for (int c = 0; c < 7; ++c) ht_sts(zr + c * 4, __uint_as_float(r[c]));
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kHtEpiThreads) : "memory");
      // the next accumulator is requested before the output phase when it is already complete (its latency hides under the phase)
      const int nbuf = (t + 1) % kHtTmemBufs;
      if (nbuf == 0) tf_par ^= 1;
      bool next_issued = false;
      if (t + 1 < n_tiles && ptx::mbar_test(ptx::smem_u32(&tfull[nbuf]), tf_par)) {
        ptx::tc_fence_after();
        ht_tmem_ld8(t_addr + nbuf * 32, r);
        next_issued = true;
      }
      // outputs of the flattened positions [128 t - 6, 128 t + 122): every Z row they read is in the ring now
      xo += 128; while (xo >= P) { xo -= P; ++s; }
      zb += 128; if (zb >= kHtZRing) zb -= kHtZRing;
      if (xo >= 0 && xo < W && co < Cout && s < n_rows) {
        long long rb;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(rb) : "r"(rt_s + (uint32_t)s * 8) : "memory");
        if (rb >= 0) {                                            // not a row that mixes two frame segments
          float v = bias;
#pragma unroll
          for (int kx = 0; kx < 7; ++kx) {
            const int i = zb + kx;
            v += ht_lds(zs_s + (uint32_t)((i >= kHtZRing ? i - kHtZRing : i) * kHtZStride + kx * Cout + co) * 4);
          }
          v = (p.act == NPVP_ACT_TANH) ? ht_tanh(v) : (p.act == NPVP_ACT_SIGMOID ? ht_sigmoid(v) : v);
          const size_t oi = (size_t)rb + plane + xo;
          if (p.out) p.out[oi] = v;
          if (p.out_u8) {   // VidReNormalize + clamp + ToPILImage in the reference's operation order (frames_to_pixels_kernel): bit-identical
            const float px = fminf(fmaxf(__fsub_rn(__fdiv_rn(v, p.pix.inv_std[co]), p.pix.inv_mean[co]), 0.0f), 1.0f);
            p.out_u8[oi] = (uint8_t)__float2uint_rz(__fmul_rn(px, 255.0f));
          }
        }
      }
      if (t + 1 < n_tiles && !next_issued) {
        ht_wait(&tfull[nbuf], tf_par, p.wait_mode);
        ptx::tc_fence_after();
        ht_tmem_ld8(t_addr + nbuf * 32, r);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, kHtTmemBufs * 32); }
}

// =============================================================================================
// 7x7 reflect-padded STEM convolution (Cin = 1 / 3 -> Cout = 32 / 64, folded BN + ReLU; ResNetAutoEncoder.py:62-66) on tcgen05
// =============================================================================================
// Same flattened-row scheme as the head above, with the roles of the taps swapped: the 7 horizontal taps x 4 padded channels are
// the K = 32 of one kernel row (the last 4 hit zero weights), so the A row of position x' for input row r is the 64 contiguous
// bytes "pixels x' .. x'+7" of the staged 16-bit row - an overlapping window no tensor map can express.  Four builder warps
// convert each input row once (fp32 planes or uint8 pixels -> 4-channel 16-bit pixels, reflection resolved in the index) and
// write its im2col block (P positions x 64 B, 64-byte swizzle) into the ring; the block serves all seven vertical taps as the
// same window one row pitch further down, exactly like the head.  N = Cout, so the accumulator IS the output: the epilogue adds the
// folded shift, applies ReLU and stores 16-bit NHWC (a thread owns a position: 32 / 64 contiguous bytes, a warp 1 / 2 KB).
// The mma.sync stem (autoencoder_kernels.cu) ran at 0.13 of the HBM roofline (212-237 us per 128 Cityscapes frames).
constexpr int kStBuilders = 8;          // warps 0, 2, 12..17 build rows (round robin); 1 / 3 issue MMAs; 4..11 epilogue
constexpr int kStThreads = 576;
constexpr int kStTmemBufs = 4;

struct StemTcParams {
  const float* x;        // fp32 frames [frames, CIN, H, W] or
  const uint8_t* x_u8;   // uint8 pixels (VidToTensor + VidNormalize in the builder, reference operation order)
  float mean[3], std[3];
  const float* w;        // fp32 [(ky, kx, ci), Cout], BN scale folded
  const float* shift;    // fp32 [Cout]
  h16* out;              // 16-bit NHWC [frames, H, W, Cout]
  int64_t total_rows;
  int H, W, P, D, R, fp16, wait_mode;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(kStThreads, 1)
conv7x7_stem_tc_kernel(const StemTcParams p) {
  constexpr int PIXB = 64;                                        // K = 32 16-bit values per position and kernel row
  extern __shared__ __align__(1024) uint8_t st_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)st_smem_raw + 1023) & ~(uintptr_t)1023);
  const int P = p.P, D = p.D, R = p.R, H = p.H, W = p.W;
  const uint32_t row_bytes = (uint32_t)P * PIXB;
  uint8_t* ring = smem;                                           // [R + D rows][P positions][64 B], swizzled
  uint8_t* wsm = ring + (size_t)(R + D) * row_bytes;              // [7 ky][COUT n][64 B], swizzled
  uint8_t* stage = wsm + 7 * COUT * PIXB;                         // [kStBuilders][P + 8 pixels][8 B]
  const int stage_pitch = (P + 8) * 8;
  uint64_t* ready = reinterpret_cast<uint64_t*>(stage + kStBuilders * stage_pitch);   // row built     [kHtRing]
  uint64_t* freeb = ready + kHtRing;                              // row consumed  [kHtRing]
  uint64_t* tfull = freeb + kHtRing;                              // [kStTmemBufs]
  uint64_t* tempty = tfull + kStTmemBufs;                         // [kStTmemBufs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kStTmemBufs);
  HtSeg* segs = reinterpret_cast<HtSeg*>(tmem_slot + 2);
  int* seg_count = reinterpret_cast<int*>(segs + kHtMaxSeg + 1);
  int64_t* rowtab = reinterpret_cast<int64_t*>(seg_count + 2);    // [n_rows] element offset of (frame, y, 0, 0) in out, -1: no output row

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t o_begin = p.total_rows * blockIdx.x / gridDim.x, o_end = p.total_rows * (blockIdx.x + 1) / gridDim.x;
  if (threadIdx.x == 0) {
    int ns = 0, s = 0;
    for (int64_t o = o_begin; o < o_end && ns < kHtMaxSeg;) {
      const int y = (int)(o % H);
      const int n = (int)min((int64_t)(H - y), o_end - o);
      segs[ns].s0 = s; segs[ns].n = n; segs[ns].f = (int)(o / H); segs[ns].y0 = y;
      s += n + 6; o += n; ++ns;
    }
    segs[ns].s0 = s; segs[ns].n = 0; segs[ns].f = 0; segs[ns].y0 = 0;
    *seg_count = ns;
    for (int i = 0; i < kHtRing; ++i) { ptx::mbar_init(&ready[i], 1); ptx::mbar_init(&freeb[i], 1); }
    for (int i = 0; i < kStTmemBufs; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 8); }
    ptx::fence_barrier_init();
  }
  // B operand: row n = output channel, K index k = kx * 4 + ci (kx = 7 and ci >= CIN are zero), one 64-byte swizzle row per (ky, n)
  {
    const uint32_t w_s = ptx::smem_u32(wsm);
    for (int i = threadIdx.x; i < 7 * COUT * 32; i += kStThreads) {
      const int k = i & 31, n = (i >> 5) % COUT, ky = i / (32 * COUT);
      const int kx = k >> 2, ci = k & 3;
      const float v = (kx < 7 && ci < CIN) ? __ldg(p.w + ((size_t)(ky * 7 + kx) * CIN + ci) * COUT + n) : 0.f;
      const int r = ky * COUT + n, c = k >> 3;
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(w_s + (uint32_t)r * PIXB + ((c ^ ((r >> 1) & 3)) << 4) + (k & 7) * 2), "h"(float_to_h16(v, p.fp16)) : "memory");
    }
  }
  ptx::fence_proxy_async();
  if (warp == 3) { ptx::tmem_alloc(tmem_slot, kStTmemBufs * COUT); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nseg = *seg_count;
  const int n_rows = segs[nseg].s0;
  const int n_tiles = n_rows > 6 ? (int)(((int64_t)(n_rows - 6) * P + 127) / 128) : 0;
  const uint32_t ring_s = ptx::smem_u32(ring);
  for (int k = 0; k < nseg; ++k)
    for (int j = threadIdx.x; j < segs[k].n + 6; j += kStThreads)
      rowtab[segs[k].s0 + j] = j < segs[k].n ? ((int64_t)segs[k].f * H + segs[k].y0 + j) * W * COUT : -1;
  __syncthreads();

  const int builder = warp == 0 ? 0 : (warp == 2 ? 1 : (warp >= 12 ? warp - 10 : -1));
  if (builder >= 0) {
    // ---------------- row builders: stream row s = builder, builder + kStBuilders, ... ----------------
    // The pixels of a builder's NEXT row are requested before it writes the current row's im2col block, so the DRAM latency of
    // the fp32 / uint8 frames hides behind ~250 shared-memory instructions (first version: 4 builders, load -> convert -> build
    // in sequence = 9 k clk per row and builder, the tensor pipe waiting: 168 us per 148 frames).
    const uint32_t stg_s = ptx::smem_u32(stage) + (uint32_t)(builder * stage_pitch);
    constexpr int kLoads = (256 + 8 + 31) / 32;                   // staged pixels per lane, P <= 256
    float raw[kLoads][CIN];
    int seg = 0;
    auto load_row = [&](int s) {
      while (s >= segs[seg + 1].s0) ++seg;
      const int f = segs[seg].f;
      const int yy = ht_reflect(segs[seg].y0 - 3 + (s - segs[seg].s0), H);
      // staged pixel i is image column reflect(i - 3); columns past the right padding repeat the last one (they only meet zero weights)
#pragma unroll
      for (int u = 0; u < kLoads; ++u) {
        const int i = lane + 32 * u;
        if (i < P + 8) {
          const int xx = ht_reflect(min(i - 3, W + 2), W);
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) {
            const size_t off = (((size_t)f * CIN + ci) * H + yy) * W + xx;
            raw[u][ci] = p.x_u8 ? (float)__ldg(p.x_u8 + off) : __ldg(p.x + off);
          }
        }
      }
    };
    if (builder < n_rows) load_row(builder);
    for (int s = builder; s < n_rows; s += kStBuilders) {
      const int slot = s % R, use = s / R;
      __syncwarp();                                               // the previous row's im2col reads of the staging row are done
#pragma unroll
      for (int u = 0; u < kLoads; ++u) {
        const int i = lane + 32 * u;
        if (i < P + 8) {
          float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci)     // uint8 pixels: VidToTensor + VidNormalize in the reference's operation order
            v[ci] = p.x_u8 ? __fdiv_rn(__fsub_rn(__fdiv_rn(raw[u][ci], 255.0f), p.mean[ci]), p.std[ci]) : raw[u][ci];
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(stg_s + (uint32_t)i * 8), "r"(pack_h16x2(v[0], v[1], p.fp16)), "r"(pack_h16x2(v[2], v[3], p.fp16)) : "memory");
        }
      }
      __syncwarp();
      if (s + kStBuilders < n_rows) load_row(s + kStBuilders);
      if (use > 0) ht_wait(&freeb[slot], (use - 1) & 1, p.wait_mode);
      // im2col block: 16-byte chunk c of position x' = staged pixels x' + 2c, x' + 2c + 1
      for (int rep = 0; rep < (slot < D ? 2 : 1); ++rep) {
        const uint32_t row = ring_s + (uint32_t)(rep ? R + slot : slot) * row_bytes;
        for (int idx = lane; idx < P * 4; idx += 32) {
          const int xp = idx >> 2, c = idx & 3;
          uint32_t a0, a1, a2, a3;
          const uint32_t src = stg_s + (uint32_t)(xp + 2 * c) * 8;
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(src) : "memory");
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a2), "=r"(a3) : "r"(src + 8) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (uint32_t)xp * PIXB + ((c ^ ((xp >> 1) & 3)) << 4)), "r"(a0), "r"(a1), "r"(a2), "r"(a3) : "memory");
        }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ht_mbar_arrive(&ready[slot]);
    }
  } else if (warp == 1 || warp == 3) {
    // ---------------- two MMA issuers, alternate tiles: 7 taps x 2 MMAs (M 128, N COUT, K 16) ----------------
    if (ptx::elect_one()) {
      const int me = warp >> 1;
      const uint32_t idesc = ht_idesc(128, COUT, p.fp16);
      const uint32_t hi = (uint32_t)(ht_desc<PIXB>(0) >> 32);
      const uint32_t a_base = (uint32_t)ht_desc<PIXB>(ring_s), b_base = (uint32_t)ht_desc<PIXB>(ptx::smem_u32(wsm));
      const uint32_t row16 = row_bytes >> 4;
      int rows_ready = 0, rdy_slot = 0;
      uint32_t rdy_par = 0;
      int r0 = 0, off = 0, slot0 = 0;
      if (me) { off = 128; while (off >= P) { off -= P; ++r0; if (++slot0 == R) slot0 = 0; } }
      uint32_t te_par = 0;
      for (int t = me; t < n_tiles; t += 2) {
        int r_last = r0 + 6, o2 = off + 127;
        while (o2 >= P) { o2 -= P; ++r_last; }
        r_last = min(r_last, n_rows - 1);
        while (rows_ready <= r_last) {
          ht_wait(&ready[rdy_slot], rdy_par, p.wait_mode);
          ++rows_ready;
          if (++rdy_slot == R) { rdy_slot = 0; rdy_par ^= 1; }
        }
        const int buf = t % kStTmemBufs;
        if (t >= kStTmemBufs) ht_wait(&tempty[buf], te_par, p.wait_mode);
        if (buf >= kStTmemBufs - 2 && t >= kStTmemBufs) te_par ^= 1;
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * COUT;
        const uint32_t a_off = a_base + (uint32_t)off * (PIXB / 16);
        int slot = slot0;
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
          const uint32_t a_lo = a_off + (uint32_t)slot * row16, b_lo = b_base + ky * (COUT * PIXB / 16);
          if (++slot == R) slot = 0;
          ht_umma_tap<2>(d_tmem, a_lo, b_lo, hi, idesc, ky ? 1u : 0u);
        }
        ptx::umma_commit(&tfull[buf]);
        off += 256;
        while (off >= P) { off -= P; ++r0; if (++slot0 == R) slot0 = 0; }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ---------------- epilogue: 8 warps = 4 TMEM lane quadrants x 2 column halves; a thread owns one position ----------------
    constexpr int CW = COUT / 2;                                  // channels per warp: 16 / 32
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int m = quad * 32 + lane;
    float sh[CW];
#pragma unroll
    for (int c = 0; c < CW; ++c) sh[c] = __ldg(p.shift + half * CW + c);
    const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + half * CW;
    const uint32_t rt_s = ptx::smem_u32(rowtab);
    int s = 0, xo = m - 128;                                      // stream row / column of this thread's position, one tile behind
    int fr_row = 0, fr_off = 0, fr_slot = 0;
    uint32_t tf_par = 0;
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t % kStTmemBufs;
      if (t && buf == 0) tf_par ^= 1;
      ht_wait(&tfull[buf], tf_par, p.wait_mode);
      ptx::tc_fence_after();
      uint32_t r[CW];
      if (CW == 16) ht_tmem_ld16(t_addr + buf * COUT, reinterpret_cast<uint32_t(&)[16]>(r));
      else ptx::tmem_ld_32x32(t_addr + buf * COUT, reinterpret_cast<uint32_t(&)[32]>(r));
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ht_mbar_arrive(&tempty[buf]);
      if (threadIdx.x == 128) {                                   // every MMA of the tiles <= t has completed: rows in front of tile t + 1 are free
        fr_off += 128;
        while (fr_off >= P) { fr_off -= P; if (fr_row < n_rows) ht_mbar_arrive(&freeb[fr_slot]); ++fr_row; if (++fr_slot == R) fr_slot = 0; }
      }
      xo += 128; while (xo >= P) { xo -= P; ++s; }
      if (xo < W && s < n_rows) {
        long long rb;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(rb) : "r"(rt_s + (uint32_t)s * 8) : "memory");
        if (rb >= 0) {
          uint32_t pk[CW / 2];
#pragma unroll
          for (int c = 0; c < CW / 2; ++c)
            pk[c] = pack_h16x2(fmaxf(__uint_as_float(r[2 * c]) + sh[2 * c], 0.f), fmaxf(__uint_as_float(r[2 * c + 1]) + sh[2 * c + 1], 0.f), p.fp16);
          h16* dst = p.out + (size_t)rb + (size_t)xo * COUT + half * CW;
#pragma unroll
          for (int c = 0; c < CW / 8; ++c)
            *reinterpret_cast<uint4*>(dst + 8 * c) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 3) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, kStTmemBufs * COUT); }
}

int g_stem_tc = 1;    // npvp_set_option("stem_tc", 0): always the mma.sync stem kernel (A/B switch)
int g_head_tc_wait = 1;
int g_head_tc = 1;    // npvp_set_option("head_tc", 0): always the mma.sync head kernel (A/B switch)

}  // namespace

void npvp_head_tc_enable(int v) { g_head_tc = v; }
void npvp_head_tc_wait(int v) { g_head_tc_wait = v ? 1 : 0; }

// Returns NPVP_OK after launching, or 1 when the geometry is not covered (the caller runs the mma.sync kernel).
int npvp_conv7x7_head_tc_try(const void* x, const void* w, const float* bias, float* out, int64_t frames, int Cin, int Cout, int H, int W,
                             int act, int fp16, uint8_t* out_u8, const float* pix_inv_std, const float* pix_inv_mean, cudaStream_t st) {
  if (!g_head_tc || (Cin != 32 && Cin != 64) || W + 6 > 256 || H < 4 || W < 4 || frames * (int64_t)H >= (1ll << 31)) return 1;
  static int num_sms = 0, max_smem = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  const int PIXB = Cin * 2;
  const int P = (W + 6 + 7) / 8 * 8;
  const int D = (P - 8 + 127) / P;
  const int fixed = 7 * 32 * PIXB + kHtZRing * kHtZStride * 4 + 16 + (3 * kHtRing + 2 * kHtTmemBufs) * 8 + 16 + (kHtMaxSeg + 1) * 16 + 16 + kHtMaxRows * 8 + 1024;
  const int R = min(kHtRing, (max_smem - fixed) / (P * PIXB) - D);
  if (D > kHtMaxDup || R < D + 8) return 1;
  const int smem = fixed + (R + D) * P * PIXB;

  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { npvp_set_error("cuTensorMapEncodeTiled entry point unavailable"); return NPVP_ERR_CUDA; }
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)(frames * H)};
  cuuint64_t strides[2] = {(cuuint64_t)PIXB, (cuuint64_t)W * PIXB};
  cuuint32_t box[3] = {(cuuint32_t)Cin, (cuuint32_t)P, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, Cin == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { npvp_set_error("conv7x7_head: cuTensorMapEncodeTiled failed (%d): frames=%lld H=%d W=%d Cin=%d", (int)r, (long long)frames, H, W, Cin); return NPVP_ERR_CUDA; }

  HeadTcParams p = {};
  p.w = (const h16*)w; p.bias = bias; p.out = out; p.out_u8 = out_u8;
  p.total_rows = frames * H;
  p.wait_mode = g_head_tc_wait;
  p.H = H; p.W = W; p.P = P; p.D = D; p.R = R; p.Cout = Cout; p.NT = (7 * Cout + 7) / 8; p.act = act; p.fp16 = fp16;
  if (out_u8) for (int c = 0; c < Cout && c < 3; ++c) { p.pix.inv_std[c] = pix_inv_std[c]; p.pix.inv_mean[c] = pix_inv_mean[c]; }
  // one CTA per SM when there is enough work (>= 16 output rows each); more CTAs only to bound the frame segments of a CTA
  int64_t grid = min((int64_t)num_sms, max((int64_t)1, p.total_rows / 16));
  grid = max(grid, (frames + kHtMaxSeg - 5) / (kHtMaxSeg - 4));
  grid = max(grid, (p.total_rows + 1023) / 1024);               // stream rows per CTA <= 1024 + 6 per segment <= kHtMaxRows
  if (grid > num_sms) grid = (grid + num_sms - 1) / num_sms * num_sms;   // whole waves
  auto kern = Cin == 64 ? conv7x7_head_tc_kernel<64> : conv7x7_head_tc_kernel<32>;
  static int attr[2] = {0, 0};
  if (attr[Cin == 64] < smem) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) { npvp_set_error("conv7x7_head_tc: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(err)); return NPVP_ERR_CUDA; }
    attr[Cin == 64] = smem;
  }
  kern<<<(unsigned)grid, kHtThreads, smem, st>>>(tm, p);
  NPVP_LAUNCH_CHECK("conv7x7_head_tc_kernel");
  return NPVP_OK;
}

void npvp_stem_tc_enable(int v) { g_stem_tc = v; }

// Returns NPVP_OK after launching, or 1 when the geometry is not covered (the caller runs the mma.sync kernel).
int npvp_conv7x7_stem_tc_try(const float* x, const uint8_t* x_u8, const float* mean, const float* stdv, const float* w, const float* shift, void* out,
                             int64_t frames, int Cin, int Cout, int H, int W, int fp16, cudaStream_t st) {
  if (!g_stem_tc || (Cin != 1 && Cin != 3) || (Cout != 32 && Cout != 64) || W + 6 > 256 || H < 4 || W < 4 || frames * (int64_t)H >= (1ll << 31)) return 1;
  static int num_sms = 0, max_smem = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  const int P = (W + 6 + 7) / 8 * 8;
  const int D = (P - 8 + 127) / P;
  const int fixed = 7 * Cout * 64 + kStBuilders * (P + 8) * 8 + (2 * kHtRing + 2 * kStTmemBufs) * 8 + 16 + (kHtMaxSeg + 1) * 16 + 16 + kHtMaxRows * 8 + 1024;
  const int R = min(kHtRing, (max_smem - fixed) / (P * 64) - D);
  if (D > kHtMaxDup || R < D + 8) return 1;
  const int smem = fixed + (R + D) * P * 64;
  StemTcParams p = {};
  p.x = x; p.x_u8 = x_u8; p.w = w; p.shift = shift; p.out = (h16*)out;
  if (x_u8) for (int c = 0; c < Cin; ++c) { p.mean[c] = mean[c]; p.std[c] = stdv[c]; }
  p.total_rows = frames * H;
  p.H = H; p.W = W; p.P = P; p.D = D; p.R = R; p.fp16 = fp16; p.wait_mode = g_head_tc_wait;
  int64_t grid = min((int64_t)num_sms, max((int64_t)1, p.total_rows / 16));
  grid = max(grid, (frames + kHtMaxSeg - 5) / (kHtMaxSeg - 4));
  grid = max(grid, (p.total_rows + 1023) / 1024);
  if (grid > num_sms) grid = (grid + num_sms - 1) / num_sms * num_sms;
  void (*kern)(const StemTcParams) = Cin == 1 ? (Cout == 32 ? conv7x7_stem_tc_kernel<1, 32> : conv7x7_stem_tc_kernel<1, 64>)
                                              : (Cout == 32 ? conv7x7_stem_tc_kernel<3, 32> : conv7x7_stem_tc_kernel<3, 64>);
  static int attr[4] = {0, 0, 0, 0};
  const int ki = (Cin == 3) * 2 + (Cout == 64);
  if (attr[ki] < smem) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) { npvp_set_error("conv7x7_stem_tc: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(err)); return NPVP_ERR_CUDA; }
    attr[ki] = smem;
  }
  kern<<<(unsigned)grid, kStThreads, smem, st>>>(p);
  NPVP_LAUNCH_CHECK("conv7x7_stem_tc_kernel");
  return NPVP_OK;
}
