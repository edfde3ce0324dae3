// Conv-FFN middle of MlpDWBN (models/VidHRFormer.py:374-392, the part between fc1 and fc2) in ONE pass and in packed half
// arithmetic:   out = GELU(LN2(dw3x3(GELU(LN1(h1))) + b))   per frame, LayerNorm over the whole (2048, 8, 8) frame with a
// per-element affine.
//
// Why this shape (r01 -> r02).  The two-kernel fp32 path (ffn_dwconv2 + ffn_norm2) moved 8 B per element and was bound by the
// FMA pipe, not by HBM: ~54 FMA-pipe cycles per channel pair and pixel in packed fp32 (FFMA2 retires 64 lanes per clock)
// = 64 us per 640 frames before any memory time.  Here
//   * all element-wise math is half2 (HFMA2: 2 values per lane per issue slot at full rate), statistics stay fp32 / fp64;
//   * GELU(x) = max(x, 0) - |x| 2^P2(|x|) with a quadratic P2 (all coefficients negative: no overflow hazards), 7 issue slots
//     per pair (3 HFMA2 + 2 MUFU.EX2 + PRMT + HMNMX2); max |error| 1.4e-3 at |x| ~ 2.7 (half an ulp of the result is 1e-3);
//   * h1 is read once and the result written once (4 B per element): a warp owns 64 channels (one pair per lane) of one
//     frame; its 64 px x 64 ch tile (8 KB of shared memory) is filled by 16-byte cp.async one frame ahead, overwritten in
//     place by the depthwise-conv output, and kept there while the LN2 statistics of the frame are gathered from the 32
//     warps (on 32 SMs) that hold the frame's other channels;
//   * the LayerNorm parameters of a block's 64 channels (2 norms x 64 px x 64 ch x (w, b) half = 32 KB) stay in shared
//     memory for the whole launch; 4 warps (4 different frame streams, two tiles each) share them, two blocks per SM;
//   * the statistics exchange goes through L2: each warp publishes (sum, sumsq) with one 8-byte store whose value is its
//     own flag (the table rests at an all-ones bit pattern no arithmetic produces), polls the frame's 256-byte line, and
//     the LAST of the 32 readers of a frame (atomic counter) puts the line back to rest, so the table is reusable by the
//     next launch without a memset.  All blocks of the grid are co-resident by construction (host sizes the grid from the
//     occupancy query), so the spin cannot deadlock; it is bounded and traps rather than hangs.
//   * the statistics of LN1 come from the fc1 GEMM epilogue's partial sums (npvp_epilogue_t.frame_stats) and are reduced
//     here, which removes the separate finalize launch.
#include "common.cuh"

namespace {

constexpr int kTok = 64;
constexpr int kCh = 2048;                  // hidden width (Spatial_FFN_hidden_ratio 4 x embed 512: NPVP's only one)
constexpr int kChunkCh = 64;               // channels per warp / block (one half2 pair per lane)
constexpr int kChunks = kCh / kChunkCh;    // 32 = number of partial statistics per frame = warp size
// Two schedules (template parameters TILES, WARPS of the kernel):
//   <2, 4> (default): two 8 KB tiles per warp, per-warp software pipeline (below), 2 x 4 warps per SM
//   <1, 8>          : one tile per warp, ph1 and ph2 of a frame back to back, 2 x 8 warps per SM whose frame periods are
//                     staggered at kernel entry so that the rendezvous of one warp is covered by the others of its scheduler
constexpr int kParamBytes = 2 * kTok * 32 * 8;            // [2 norms][64 px][32 pairs] (w2, b2) half2 x 2
constexpr int kTileBytes = kTok * 32 * 4;                 // [64 px][32 pairs] half2 = 8 KB
constexpr int smem_bytes(int tiles, int warps) { return kParamBytes + warps * tiles * kTileBytes; }   // 96 KB -> two blocks per SM
constexpr float kEps = 1e-5f;

typedef uint32_t h2;                                       // packed half2 bits

__device__ __forceinline__ h2 hfma2(h2 a, h2 b, h2 c) { h2 d; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ h2 hmul2(h2 a, h2 b) { h2 d; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ h2 habs2(h2 a) { h2 d; asm("abs.f16x2 %0, %1;" : "=r"(d) : "r"(a)); return d; }
__device__ __forceinline__ h2 hneg2(h2 a) { h2 d; asm("neg.f16x2 %0, %1;" : "=r"(d) : "r"(a)); return d; }
__device__ __forceinline__ h2 hmax2_0(h2 a) { h2 d; asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(0u)); return d; }
__device__ __forceinline__ h2 hex2(h2 a) { h2 d; asm("ex2.approx.f16x2 %0, %1;" : "=r"(d) : "r"(a)); return d; }
__device__ __forceinline__ h2 h2_bcast(float v) {
  const __half2 t = __float2half2_rn(v);
  return *reinterpret_cast<const h2*>(&t);
}
__device__ __forceinline__ float2 h2_to_f2(h2 a) { return __half22float2(*reinterpret_cast<const __half2*>(&a)); }

// log2(0.5 erfc(u / sqrt 2)) ~ P2(u) = c2 u^2 + c1 u + c0, minimax on the GELU error u 2^P (4.4e-4 in exact arithmetic; 1.4e-3
// max / 1.7e-4 mean once every step is rounded to half - the result's own rounding, half an ulp = 1e-3 at |x| ~ 3, dominates,
// which is why the cubic, 1.0e-3 max / 0.9e-4 mean, buys nothing for its extra HFMA2).  All coefficients are negative, so
// for large |x| the Horner chain runs monotonically to -inf and 2^P to 0: no clamp, no inf - inf.
// Cost per PAIR: 3 HFMA2 + 2 MUFU.EX2.F16 + PRMT + HMNMX2 = 7 issue slots.  Measured on B200 (tools/ubench/hfma2.cu): HFMA2
// issues at 0.5 / clk / sub-partition - two values per lane at half the instruction rate, i.e. the SAME 128 FMA lanes per clock
// per SM as scalar FFMA or packed FFMA2 - and MUFU.EX2.F16 at 1 / 8 clk, so every HFMA2 saved is 2 pipe cycles: the
// kernel's ceiling is the FMA pipe (~23 FMA-pipe instructions per pair and pixel), not issue slots and not HBM.
constexpr uint32_t kC2 = 0xB8A9B8A9u;   // -0.58239
constexpr uint32_t kC1 = 0xBC35BC35u;   // -1.05181
constexpr uint32_t kC0 = 0xBC16BC16u;   // -1.02189
__device__ __forceinline__ h2 gelu_h2(h2 x) {
  const h2 u = habs2(x);
  h2 p = hfma2(u, kC2, kC1);
  p = hfma2(p, u, kC0);
  return hfma2(hneg2(u), hex2(p), hmax2_0(x));
}

// the same GELU on N values, the N dependency chains advanced one step at a time
template <int N>
__device__ __forceinline__ void gelu_h2_xN(h2 (&x)[N]) {
  h2 p[N];
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = hfma2(habs2(x[i]), kC2, kC1);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = hfma2(p[i], habs2(x[i]), kC0);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = hex2(p[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = hfma2(hneg2(habs2(x[i])), p[i], hmax2_0(x[i]));
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// (sum, sumsq) partials held one per lane -> (rstd, -mean * rstd) of a frame of 1 / inv_n elements.  The 32 partials are
// summed in fp32 (fixed butterfly order: deterministic); only E[x^2] - mean^2 is formed in fp64 (cancellation).
__device__ __forceinline__ void stats_from_partials(float ps, float pq, double inv_n, float& rstd, float& nmr) {
  const float s = warp_sum(ps), q = warp_sum(pq);
  const double mean = (double)s * inv_n;
  const float var = (float)fmax((double)q * inv_n - mean * mean, 0.0);
  rstd = rsqrtf(var + kEps);
  nmr = -(float)mean * rstd;
}

template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_dyn(int n) {           // n is a compile-time value after unrolling
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    case 6: cp_async_wait<6>(); break;
    default: cp_async_wait<7>(); break;
  }
}

// Per-warp software pipeline over its frames f_0, f_1, ... (two 8 KB tiles, tile k & 1 belongs to frame f_k):
//     ph1(f_0) | ph1(f_1) ph2(f_0) | ph1(f_2) ph2(f_1) | ...
// ph1 = GELU(LN1) + depthwise conv into the tile (in place over h1) + publish the partial statistics of y2;
// ph2 = gather the frame's 32 partials, GELU(LN2), store - and refill each finished image row of the tile with h1 of the
// frame two steps ahead (16-byte cp.async, one commit group per image row, consumed row by row in the next-but-one ph1).
// Between publishing the statistics of a frame and needing everybody's, a warp computes a whole ph1: the exchange latency
// and the skew between the 32 SMs that share a frame are hidden by construction, not by occupancy.  (r02 ncu of the
// unpipelined version with 16 warps per SM: all warps of an SM reach the rendezvous together, 37% of the stall samples in the
// poll loop, 23% on the partial-statistics load at the head of a frame, 38% of the issue slots used.)
template <int TILES, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2)
ffn_mid16_kernel(const __half* __restrict__ h1, const float2* __restrict__ part1, const uint2* __restrict__ ln_wb,
                 const __half* __restrict__ dw_w, const __half* __restrict__ dw_b, __half* __restrict__ out,
                 float2* __restrict__ xch, unsigned int* __restrict__ cnt, int frames) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint2* s_wb = reinterpret_cast<uint2*>(smem);                               // [2][64][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = WARPS, kThreads = WARPS * 32;
  h2* s_tiles = reinterpret_cast<h2*>(smem + kParamBytes) + warp * (TILES * kTok * 32);   // [TILES][64 px][32 pairs] of this warp
  const int chunk = blockIdx.x;
  const int c0 = chunk * kChunkCh + 2 * lane;

  // stage this block's LayerNorm parameters: global [2][64][kCh / 2] uint2 -> shared [2][64][32] uint2 (256 B rows)
  {
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_wb);
    for (int i = threadIdx.x; i < 2 * kTok * 16; i += kThreads) {             // 16-byte pieces: 16 per 256-byte row
      const int row = i >> 4, piece = i & 15;
      cp_async16(sbase + (uint32_t)(row * 256 + piece * 16), reinterpret_cast<const uint8_t*>(ln_wb + (size_t)row * (kCh / 2) + chunk * 32) + piece * 16);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  h2 w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w[k] = __ldg(reinterpret_cast<const uint32_t*>(dw_w + (size_t)k * kCh + c0));
  const h2 bias = __ldg(reinterpret_cast<const uint32_t*>(dw_b + c0));

  constexpr double inv_n = 1.0 / ((double)kTok * (double)kCh);
  const int stride = gridDim.y * kWarps;
  const uint32_t tiles_u32 = (uint32_t)__cvta_generic_to_shared(s_tiles);
  // one image row (8 px x 128 B) of frame f -> tile t: 16-byte cp.async, a warp instruction covers 4 pixels
  const int pf_px = lane >> 3, pf_off = (lane & 7) * 16;
  auto prefetch_row = [&](int f, int t, int r) {
    const uint8_t* g = reinterpret_cast<const uint8_t*>(h1 + ((size_t)f * kTok + r * 8 + pf_px) * kCh + chunk * kChunkCh) + pf_off;
    const uint32_t d = tiles_u32 + (uint32_t)(t * (kTok * 128) + (r * 8 + pf_px) * 128 + pf_off);
    cp_async16(d, g);
    cp_async16(d + 4 * 128, g + (size_t)4 * kCh * 2);
  };
  const int f0 = blockIdx.y * kWarps + warp;
#pragma unroll
  for (int r = 0; r < 8; ++r) {                                       // first frame -> tile 0: one commit group per image row
    if (f0 < frames) prefetch_row(f0, 0, r);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  float2 p1 = make_float2(0.f, 0.f);                                  // fc1's partial statistics of the frame ph1 runs next
  if (f0 < frames) p1 = __ldg(part1 + (size_t)f0 * kChunks + lane);
  asm volatile("cp.async.wait_group 8;" ::: "memory");               // the parameters (everything older than the 8 row groups)
  __syncthreads();
  if (f0 >= frames) return;

  // ---- ph1 of frame f in tile t (requested one image row per commit group by the previous ph2 / the prologue) ----
  auto phase1 = [&](int f, int t, float2 pp) {
    float rstd1, nmr1;
    stats_from_partials(pp.x, pp.y, inv_n, rstd1, nmr1);
    const h2 rs1 = h2_bcast(rstd1), nm1 = h2_bcast(nmr1);
    h2* tile = s_tiles + t * (kTok * 32);
    // two image rows per step (16 independent GELU chains, then 16 conv chains): with two warps per scheduler the kernel is
    // bound by per-warp dependency latency (r02 ncu: 1.1 `wait` stall cycles per issue with 8 chains), not by issue slots
    h2 g[4][8];                                                      // GELU(LN1) rows, row r lives in g[r & 3]
    f32x2 s2 = pk2(0.f, 0.f), q2 = pk2(0.f, 0.f);
    auto conv_rows = [&](int o0, int n) {                            // conv output rows o0 .. o0 + n - 1 (n <= 2), tap-major: 8 n independent chains
      h2 acc[2][8];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int x = 0; x < 8; ++x) acc[j][x] = bias;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (j >= n) continue;
            const int iy = o0 + j + ky - 1;
            if (iy < 0 || iy > 7) continue;
#pragma unroll
            for (int x = 0; x < 8; ++x) {
              const int ix = x + kx - 1;
              if (ix >= 0 && ix <= 7) acc[j][x] = hfma2(g[iy & 3][ix], w[ky * 3 + kx], acc[j][x]);
            }
          }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (j >= n) continue;
#pragma unroll
        for (int x = 0; x < 8; ++x) {                                // in place over h1 (row o0 + j <= the rows already consumed)
          tile[((o0 + j) * 8 + x) * 32 + lane] = acc[j][x];
          const float2 a = h2_to_f2(acc[j][x]);
          const f32x2 a2 = pk2(a.x, a.y);
          s2 = add2(s2, a2);
          q2 = fma2(a2, a2, q2);
        }
      }
    };
#pragma unroll
    for (int rp = 0; rp < 4; ++rp) {
      cp_async_wait_dyn(6 - 2 * rp);                                 // image rows 0 .. 2 rp + 1 of this frame have landed
      __syncwarp();
      h2 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = hfma2(tile[(rp * 16 + i) * 32 + lane], rs1, nm1);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint2 wb = s_wb[(rp * 16 + i) * 32 + lane];
        v[i] = hfma2(v[i], wb.x, wb.y);
      }
      gelu_h2_xN<16>(v);
#pragma unroll
      for (int i = 0; i < 16; ++i) g[(2 * rp + (i >> 3)) & 3][i & 7] = v[i];
      if (rp > 0) conv_rows(2 * rp - 1, 2);
      else conv_rows(0, 1);
    }
    conv_rows(7, 1);
    float s_lo, s_hi, q_lo, q_hi;
    upk2(s2, s_lo, s_hi);
    upk2(q2, q_lo, q_hi);
    const float ps = warp_sum(s_lo + s_hi), pq = warp_sum(q_lo + q_hi);
    if (lane == 0) asm volatile("st.relaxed.gpu.global.v2.f32 [%0], {%1, %2};" ::"l"(xch + (size_t)f * kChunks + chunk), "f"(ps), "f"(pq) : "memory");
  };

  // ---- ph2 of frame f in tile t; each finished image row is refilled with h1 of frame f_pre (if any) ----
  auto phase2 = [&](int f, int t, int f_pre) {
    float vx, vy;
    {
      const float2* line = xch + (size_t)f * kChunks + lane;
      uint32_t spins = 0;
      long long t0 = 0;
      for (;;) {
        asm volatile("ld.relaxed.gpu.global.v2.f32 {%0, %1}, [%2];" : "=f"(vx), "=f"(vy) : "l"(line) : "memory");
        if (__all_sync(0xffffffffu, __float_as_uint(vx) != 0xffffffffu)) break;
        __nanosleep(64);
        if ((++spins & 0xFFu) == 0) {                                // bounded: a missing peer traps instead of hanging the GPU
          const long long now = clock64();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 4000000000LL) __trap();
        }
      }
    }
    unsigned int readers = 0;                                        // readers of this frame's line before this warp (used after the row loop)
    if (lane == 0) readers = atomicAdd(cnt + f, 1u);
    float rstd2, nmr2;
    stats_from_partials(vx, vy, inv_n, rstd2, nmr2);
    const h2 rs2 = h2_bcast(rstd2), nm2 = h2_bcast(nmr2);
    const h2* tile = s_tiles + t * (kTok * 32) + lane;
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (size_t)f * kTok * kCh + c0);
    const uint2* wb2 = s_wb + kTok * 32 + lane;
#pragma unroll 1
    for (int r = 0; r < 8; r += 2) {                                 // two image rows (16 independent chains) per iteration
      h2 y[16];
#pragma unroll
      for (int x = 0; x < 16; ++x) y[x] = hfma2(tile[(r * 8 + x) * 32], rs2, nm2);
#pragma unroll
      for (int x = 0; x < 16; ++x) {
        const uint2 wb = wb2[(r * 8 + x) * 32];
        y[x] = hfma2(y[x], wb.x, wb.y);
      }
      gelu_h2_xN<16>(y);
#pragma unroll
      for (int x = 0; x < 16; ++x) dst[(size_t)(r * 8 + x) * (kCh / 2)] = y[x];
      __syncwarp();                                                  // every lane has read these rows of y2
      if (f_pre < frames) prefetch_row(f_pre, t, r);
      asm volatile("cp.async.commit_group;" ::: "memory");          // always: ph1 counts on 8 groups per ph2
      if (f_pre < frames) prefetch_row(f_pre, t, r + 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (__shfl_sync(0xffffffffu, readers, 0) == kChunks - 1) {       // the last reader of the frame puts its line back to rest
      asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" ::"l"(xch + (size_t)f * kChunks + lane), "r"(0xffffffffu), "r"(0xffffffffu) : "memory");
      if (lane == 0) cnt[f] = 0u;
    }
  };

  if (TILES == 1) {
    // single tile: ph1(f) | rendezvous | ph2(f) (refilling the tile with the next frame).  The warps of a block start their
    // first frame one eighth of a frame period apart, so the rendezvous latency of one is hidden by its scheduler's others.
    __nanosleep(1200u * (unsigned)warp);
#pragma unroll 1
    for (int f = f0; f < frames; f += stride) {
      const float2 p1_now = p1;
      if (f + stride < frames) p1 = __ldg(part1 + (size_t)(f + stride) * kChunks + lane);
      phase1(f, 0, p1_now);
      phase2(f, 0, f + stride);
    }
    cp_async_wait<0>();
    return;
  }
  int t = 0;
#pragma unroll 1
  for (int f = f0;; f += stride, t ^= 1) {                            // f: frame of this step's ph1 (tile t); ph2 of the previous frame
    const bool live = f < frames;
    if (live) {
      const float2 p1_now = p1;
      if (f + stride < frames) p1 = __ldg(part1 + (size_t)(f + stride) * kChunks + lane);   // next step's, a whole step ahead
      phase1(f, t, p1_now);
    }
    if (f != f0) {
      phase2(f - stride, t ^ 1, f + stride);
    } else {                                                          // first step: nothing to finish yet, request the second frame
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (f + stride < frames) prefetch_row(f + stride, 1, r);
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }
    if (!live) break;
  }
  cp_async_wait<0>();
}

int g_mid16_lanes[2] = {-1, -1};   // per schedule: blockIdx.y extent that keeps the whole grid co-resident (0: does not fit)
int g_mid16_mode = 0;              // npvp_set_option("ffn_mid16_mode", m): 0 = <2 tiles, 4 warps> pipelined, 1 = <1 tile, 8 warps> staggered

template <int TILES, int WARPS>
int mid16_lanes_of() {
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaFuncSetAttribute(ffn_mid16_kernel<TILES, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(TILES, WARPS)) == cudaSuccess &&
      cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ffn_mid16_kernel<TILES, WARPS>, WARPS * 32, smem_bytes(TILES, WARPS)) == cudaSuccess)
    return sms * per_sm / kChunks;
  return 0;
}

void mid16_setup() {
  if (g_mid16_lanes[0] >= 0) return;
  g_mid16_lanes[0] = mid16_lanes_of<2, 4>();
  g_mid16_lanes[1] = mid16_lanes_of<1, 8>();
  cudaGetLastError();
}

}  // namespace

void npvp_set_ffn_mid16_mode(int m) { g_mid16_mode = m ? 1 : 0; }

extern "C" int npvp_ffn_mid16_lanes(void) {
  mid16_setup();
  return g_mid16_lanes[g_mid16_mode];
}

extern "C" int npvp_ffn_mid16(const void* h1_f16, const float* part1, const void* ln_wb_f16, const void* dw_w_f16, const void* dw_b_f16,
                              void* out_f16, float* xch, unsigned int* cnt, int64_t frames, int64_t Ch, void* stream) {
  NPVP_REQUIRE(h1_f16 && part1 && ln_wb_f16 && dw_w_f16 && dw_b_f16 && out_f16 && xch && cnt, "npvp_ffn_mid16: null pointer");
  NPVP_REQUIRE(frames > 0 && frames < (1ll << 31), "npvp_ffn_mid16: frames > 0");
  NPVP_REQUIRE(Ch == kCh, "npvp_ffn_mid16: Ch must be %d (use npvp_ffn_dwconv + npvp_ffn_norm2 otherwise)", kCh);
  NPVP_REQUIRE(h1_f16 != out_f16, "npvp_ffn_mid16: in-place operation is not supported (a frame's chunks finish at different times)");
  NPVP_REQUIRE(((uintptr_t)xch & 255) == 0 && ((uintptr_t)ln_wb_f16 & 15) == 0 && ((uintptr_t)h1_f16 & 3) == 0 && ((uintptr_t)out_f16 & 3) == 0,
               "npvp_ffn_mid16: xch must be 256-byte aligned, parameters 16-byte aligned");
  mid16_setup();
  const int mode = g_mid16_mode;
  NPVP_REQUIRE(g_mid16_lanes[mode] > 0, "npvp_ffn_mid16: the kernel does not fit on this device");
  const int warps = mode ? 8 : 4;
  const int64_t want = (frames + warps - 1) / warps;
  const int64_t lanes = want > g_mid16_lanes[mode] ? g_mid16_lanes[mode] : want;
  dim3 grid(kChunks, (unsigned)lanes);
  if (mode)
    ffn_mid16_kernel<1, 8><<<grid, 8 * 32, smem_bytes(1, 8), (cudaStream_t)stream>>>((const __half*)h1_f16, (const float2*)part1, (const uint2*)ln_wb_f16,
        (const __half*)dw_w_f16, (const __half*)dw_b_f16, (__half*)out_f16, (float2*)xch, cnt, (int)frames);
  else
    ffn_mid16_kernel<2, 4><<<grid, 4 * 32, smem_bytes(2, 4), (cudaStream_t)stream>>>((const __half*)h1_f16, (const float2*)part1, (const uint2*)ln_wb_f16,
        (const __half*)dw_w_f16, (const __half*)dw_b_f16, (__half*)out_f16, (float2*)xch, cnt, (int)frames);
  NPVP_LAUNCH_CHECK("ffn_mid16_kernel");
  return NPVP_OK;
}
