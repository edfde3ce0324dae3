// Dense contractions for the NPVP hot path on sm_100a.
//
//   npvp_gemm_bf16 : D[M,N] = A[M,K] * W[N,K]^T, bf16 operands, fp32 accumulate, fused epilogue.
//       TCGEN05 backend: TMA (cp.async.bulk.tensor, 128B swizzle) -> 3/4-stage smem ring ->
//       tcgen05.mma (cta_group::1, 128 x BN x 16, one elected issuing thread) -> TMEM accumulator ->
//       tcgen05.ld epilogue (bias / act / alpha / residuals / relu; fp32 and/or bf16 stores).
//       SIMT backend: shared-memory tiled CUDA-core kernel with the same epilogue, kept for debugging the
//       tensor path and for shapes the TMA descriptors cannot express.
//   npvp_gemm_f32  : fp32 CUDA-core GEMM for the NRMLP positional MLP (precision critical, tiny).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <string.h>


// Optional timeline instrumentation (tools/ubench/gemm_trace.cu compiles this file with -DNPVP_GEMM_TRACE): CTA 0 records
// clock64() at the pipeline hand-offs.  Never compiled into the library.
#ifdef NPVP_GEMM_TRACE
__device__ long long g_trace[8][8192];
__device__ int g_trace_n[8];
#define NPVP_TRACE(slot) do { if (blockIdx.x == 0) { int i_ = g_trace_n[slot]; if (i_ < 8192) { g_trace[slot][i_] = clock64(); g_trace_n[slot] = i_ + 1; } } } while (0)
#else
#define NPVP_TRACE(slot) do { } while (0)
#endif

// =============================================================================================
// tcgen05 GEMM kernel
// =============================================================================================
// Tile: BLOCK_M = 128 (one UMMA M), BLOCK_N = BN (UMMA N), BLOCK_K = 64 bf16 = one 128-byte swizzle row.
// smem per stage: A 128x64 bf16 (16 KB) + B BNx64 bf16.  K-major, SWIZZLE_128B (matches the TMA maps).
constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kUmmaK = 16;

// Shared-memory matrix descriptor: K-major, 128B swizzle, 8-row groups 1024 B apart (SBO), version 1 (sm_100).
// low word of the same descriptor (see umma_f16_x4)
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address, 16-byte units
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major; canonical 1)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // layout type: SWIZZLE_128B
  return d;
}

// Instruction descriptor, kind::f16: D=f32 (bit 4), A/B format at [7,10)/[10,13) (0 = f16, 1 = bf16), both K-major,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_f16kind(int M, int N, int fp16) {
  return (1u << 4) | (fp16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// =============================================================================================
// tcgen05 GEMM kernel, v2: persistent + warp-specialised + double-buffered TMEM + coalesced, specialised epilogue
// =============================================================================================
// One CTA per SM walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (n fastest, so the CTAs running at any moment
// share A row-blocks in L2 and the whole W).  Roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4-11 = epilogue (two warps per TMEM lane quadrant, each owning half of the tile's columns).
// The accumulator is double-buffered in TMEM (2 x BN columns): the epilogue of tile i overlaps the main loop of tile i+1.
// Epilogue data path: tcgen05.ld (row per thread) -> raw accumulators through an XOR-swizzled 32x32 smem transpose ->
// "coalesced" layout (a warp instruction covers 4 rows x 128 B) where bias (one float4 per lane per chunk), activation,
// alpha, residual loads and the fp32 / 16-bit stores all happen.  ACT / RES / OUT are compile-time (r01 profile: the
// first version spent >100 instructions per element on run-time flag checks and was epilogue-bound at 6 % tensor-active).
constexpr int kGemm2Threads = 384;
constexpr int kEpiWarps = 8;

// RES: 0 none, 1 res1 fp32, 2 res1 16-bit, 3 res1 + res2 16-bit, -1 run-time.   OUT: 0 16-bit, 1 fp32, 2 both, -1 run-time,
// 3 16-bit stored straight from the accumulator layout (opt-in), 4 16-bit + per-frame partial statistics (EpiParams::frame_stats).
constexpr int kOutStats16 = 4;
template <int BN>
struct Gemm2Cfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 5 : 6);
  static constexpr int kStagingBytes = kEpiWarps * 32 * 32 * 4;             // one 32x32 fp32 tile per epilogue warp
  static constexpr int kTableBytes = (BN == 256) ? 1024 : 2048;             // conv gather: [2 buffers][taps per k-block][128 rows] source pixels
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align*/ + 256 /*barriers*/ + kTableBytes;
  static constexpr uint32_t kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;          // 128 / 256 / 512
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}
// Explicit shared-space accesses for the epilogue staging tile: through a generic pointer carved out of the dynamic smem
// block the compiler emitted generic ST.E / LD.E (long-scoreboard latency, r01 profile) instead of STS / LDS.
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// Implicit-GEMM convolution: the A operand (rows = output pixels, K = (ky, kx, ci)) is gathered by two producer warps
// straight from the channels-last activations with 16-byte cp.async into the same 128B-swizzled layout TMA would
// produce (row r at r*128 B, 16-byte chunk c stored at chunk c ^ (r & 7)), so the MMA side is unchanged and no
// im2col buffer exists.  Zero / reflect / replicate padding, stride and the phase-major layout left by the transposed
// convolutions are resolved in the address math; out-of-range taps are zero-filled (cp.async src-size 0).
// (A variant with 8 lanes per 128-byte row segment - fully coalesced sectors - was measured 2.3x SLOWER: the gather is
// bound by its address arithmetic, so each thread keeps two whole rows and amortises the decode over 8 chunks.)
struct ConvGather {
  const h16* x;
  int H, W, C, KH, KW, stride, pad, pad_mode, Ho, Wo, phase_major;
};
// CONV = 2 / 3: the A operand of a stride-1 zero-padded convolution (2) or of the 2x2-neighbourhood form of a transposed
// convolution (3) is a SHIFTED WINDOW of the channels-last activations, i.e. a box of a 4-D tensor map (C, W, H, frames):
// one cp.async.bulk.tensor.4d per k-block lands the 128 pixels x 64 channels of tap (ty, tx) in the same 128B-swizzled
// layout as a dense A tile, pixels outside the frame are zero-filled by the TMA unit, and no gather warps exist
// (r01/r02 profiles: the cp.async gather kept the tensor pipe 26-50 % busy and spent 40 % of the issue slots on addresses).
// A tile is 128 consecutive pixels: whole rows of one frame (W <= 128, H*W % 128 == 0) or whole frames (128 % (H*W) == 0).
// CONV = 3 additionally (i) skips the zero blocks of the transposed convolution: with the output phases ordered
// q = (0,0), (0,1), (1,1), (1,0) along N, tap (dy,dx) feeds the CONTIGUOUS column range [lo, hi) x Cout given by
// convt_cols(), so every tap is one MMA group with N = hi - lo accumulating into TMEM columns [lo, hi) - 9 of the 16
// (phase, tap) blocks are computed, none is multiplied by zeros - and (ii) stores plain NHWC at double resolution
// (pixel (2y+py, 2x+px)), which is what makes the NEXT layer's A operand TMA-loadable.
struct TmaConv {
  int taps, KW, pad, kb_per_tap;   // tap t = (t / KW, t % KW), source shift (ty - pad, tx - pad); k-blocks of 64 channels per tap
  int HW, W;                        // pixels per frame / per row
  int cout;                         // CONV = 3: output channels per phase (N = 4 * cout)
  // Resident weights: when the live part of W fits beside >= 3 A stages (and N <= BN: one n-tile), the producer loads it ONCE
  // per CTA - tap by tap, k-block by k-block, live rows only, `wbox` rows per TMA box - and the stage ring carries A alone.
  // (r02 launch list: the last transposed conv and the 64-channel 3x3 conv spent 2.6 / 4.2 us per 128-pixel tile re-fetching
  // the same 64 / 72 KB of weights from L2 - these layers are bound by the per-SM L2 ingest, not by the tensor pipe.)
  int wres, wres_bytes, nst, wbox;
  int pf;                           // dense (CONV = 0): L2 prefetch of the A boxes a few k-blocks ahead of the ring (npvp_set_option("gemm_prefetch"))
};
__device__ __forceinline__ void convt_cols(int t, int cout, int n0, int bn, int& lo, int& hi) {
  const int a = t == 0 ? 0 : (t == 1 ? cout : 2 * cout);
  const int b = (t == 0 || t == 2) ? 4 * cout : 3 * cout;
  lo = max(a, n0) - n0;
  hi = min(b, n0 + bn) - n0;
}

constexpr int kGatherThreads = 128;    // warps 2, 3 and (conv kernels only) the two extra warps 12, 13
constexpr int kGatherRowsPerPass = kGatherThreads / 8;
constexpr int kGatherLag = 2;          // cp.async groups kept in flight per thread before the stage is published

__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ int conv_reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

template <int ACT>
__device__ __forceinline__ float act_ct(float v, int act_rt) {
  if (ACT == NPVP_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == NPVP_ACT_GELU) return gelu_erf(v);
  if (ACT < 0) return apply_act(v, act_rt);
  return v;
}

__device__ __forceinline__ void add_res(float4& q, const void* res, int is16, int fp16, int64_t off) {
  if (is16) {
    const uint2 u = *reinterpret_cast<const uint2*>((const h16*)res + off);
    const float2 a = unpack_h16x2(u.x, fp16), b = unpack_h16x2(u.y, fp16);
    q.x += a.x; q.y += a.y; q.z += b.x; q.w += b.y;
  } else {
    const float4 a = *reinterpret_cast<const float4*>((const float*)res + off);
    q.x += a.x; q.y += a.y; q.z += a.z; q.w += a.w;
  }
}

// One 32x32 chunk of the epilogue in the coalesced layout (4 rows x 128 B per warp instruction).  Written branch-free:
// all eight swizzled LDS.128 and all residual loads are issued first, the math is straight-line, and the bounds checks
// only predicate the loads / stores.  (r01: with a per-row `if` the compiler emitted one BSSY/BSYNC region per row, the
// rows serialised at ~100 cycles each and the epilogue - not the 98 %-of-peak main loop - set the tile time.)
// Residual operands of one chunk as they come from memory (compile-time RES > 0): loaded a whole chunk AHEAD of their use -
// the first chunk of a tile before the wait for the accumulator, later ones while the previous chunk is being stored -
// so that their latency never sits between tcgen05.ld and the stores.  (r02 ncu of the 3x3 conv + skip at 64 channels: 45 % of
// the epilogue warps' samples were long-scoreboard stalls on these loads, 3.7 us per 128 x 64 tile.)
template <int RES>
struct ResRaw {
  uint2 a[(RES == 2 || RES == 3) ? 8 : 1];
  uint2 b[RES == 3 ? 8 : 1];
  float4 f[RES == 1 ? 8 : 1];
};
template <int RES>
__device__ __forceinline__ void load_res(ResRaw<RES>& rr, const EpiParams& ep, int sub_row, int64_t m_base, int64_t M, bool col_ok, int64_t n) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int64_t m = m_base + it * 4 + sub_row;
    const bool ok = m < M && col_ok;
    const int64_t roff = m * ep.ld_res + n;
    if (RES == 1) rr.f[it] = ok ? *reinterpret_cast<const float4*>((const float*)ep.res1 + roff) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (RES == 2 || RES == 3) rr.a[it] = ok ? *reinterpret_cast<const uint2*>((const h16*)ep.res1 + roff) : make_uint2(0u, 0u);
    if (RES == 3) rr.b[it] = ok ? *reinterpret_cast<const uint2*>((const h16*)ep.res2 + roff) : make_uint2(0u, 0u);
  }
}

// `next()` runs between the last use of `rr` and the stores: it refills rr with the residuals of the chunk that follows.
template <int ACT, int RES, int OUT, bool FP16, class Next>
__device__ __forceinline__ void epi_chunk(uint32_t stg_addr, int sub_row, int chunk, int64_t m_base, int64_t M, bool col_ok,
                                          int64_t n, const float4 b4, float alpha, float relu_floor, const EpiParams& ep,
                                          float& st_s, float& st_q, ResRaw<RES>& rr, Next&& next) {
  constexpr int fp16 = FP16 ? 1 : 0;
  float4 q[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = it * 4 + sub_row;
    q[it] = lds128(stg_addr + (uint32_t)((row * 32 + ((chunk ^ (row & 7)) << 2)) * 4));
  }
  float4 rs[8];
  if (RES != 0) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int64_t m = m_base + it * 4 + sub_row;
      rs[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (RES == 1) rs[it] = rr.f[it];
      if (RES == 2 || RES == 3) {
        const float2 a = unpack_h16x2(rr.a[it].x, fp16), b = unpack_h16x2(rr.a[it].y, fp16);
        rs[it] = make_float4(a.x, a.y, b.x, b.y);
      }
      if (RES == 3) {
        const float2 a = unpack_h16x2(rr.b[it].x, fp16), b = unpack_h16x2(rr.b[it].y, fp16);
        rs[it].x += a.x; rs[it].y += a.y; rs[it].z += b.x; rs[it].w += b.y;
      }
      if (RES < 0 && m < M && col_ok) {
        const int64_t roff = m * ep.ld_res + n;
        if (ep.res1) add_res(rs[it], ep.res1, ep.res1_bf16, fp16, roff);
        if (ep.res2) add_res(rs[it], ep.res2, ep.res2_bf16, fp16, roff);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    float4& v = q[it];
    v.x = act_ct<ACT>(v.x + b4.x, ep.act) * alpha;
    v.y = act_ct<ACT>(v.y + b4.y, ep.act) * alpha;
    v.z = act_ct<ACT>(v.z + b4.z, ep.act) * alpha;
    v.w = act_ct<ACT>(v.w + b4.w, ep.act) * alpha;
    if (RES != 0) { v.x += rs[it].x; v.y += rs[it].y; v.z += rs[it].z; v.w += rs[it].w; }
    v.x = fmaxf(v.x, relu_floor); v.y = fmaxf(v.y, relu_floor); v.z = fmaxf(v.z, relu_floor); v.w = fmaxf(v.w, relu_floor);
    if (OUT == kOutStats16) {                                        // every row / column is valid on this path (M % 64 == N % 256 == 0)
      st_s += (v.x + v.y) + (v.z + v.w);
      st_q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, st_q))));
    }
  }
  next();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int64_t m = m_base + it * 4 + sub_row;
    if (m < M && col_ok) {
      const int64_t ooff = m * ep.ld_out + n;
      if (OUT == 1 || OUT == 2 || (OUT < 0 && ep.out_f32)) *reinterpret_cast<float4*>(ep.out_f32 + ooff) = q[it];
      if (OUT == 0 || OUT == 2 || OUT == kOutStats16 || (OUT < 0 && ep.out_bf16))
        *reinterpret_cast<uint2*>(ep.out_bf16 + ooff) = make_uint2(pack_h16x2(q[it].x, q[it].y, fp16), pack_h16x2(q[it].z, q[it].w, fp16));
    }
  }
}

// DIRECT epilogue (OUT >= kDirect; the default whenever alignment allows, npvp_set_option("gemm_epi_direct", 0) = staged):
// everything happens in the accumulator layout - thread = row, a chunk = 32 consecutive columns of that row (64 B of 16-bit
// or 128 B of fp32 output, whole 32-byte sectors) - with the tcgen05.ld of chunk i+1 in flight while chunk i is converted
// and stored, residuals loaded a chunk ahead, and NO shared-memory transpose.
// Why (r02 timeline, tools/ubench/gemm_trace.cu): shared memory is the contended resource of this kernel.  Per 64-wide k-block
// a 128 x 256 tile has TMA write 48 KB and the tensor core read 48 KB - 750 clk of the 128 B/clk port against 512 clk of MMA -
// and the staged epilogue added a 32 KB write + 32 KB read per 32 x 256 columns on top: 5900 clk per tile (7800 in the 2-CTA
// kernel) against 4400 for this path.  alpha == 1 and no post-ReLU are required (host-side choice): the per-element work is
// bias add, activation, residual adds, rounding.
constexpr int kDirect = 8;
template <int RES>
struct ResD {                                    // residuals of one chunk of one row, as loaded (32-byte sector loads)
  uint32_t a[(RES == 2 || RES == 3) ? 16 : 1];   // 32 x 16-bit
  uint32_t b[RES == 3 ? 16 : 1];
  uint32_t f[RES == 1 ? 32 : 1];                 // 32 x fp32
};
template <int RES>
__device__ __forceinline__ void load_resd(ResD<RES>& rs, const EpiParams& ep, int64_t m, int64_t n0, bool ok) {
  const int64_t off = m * ep.ld_res + n0;
  if (!ok) {                                     // (values of masked rows are never stored)
    return;
  }
  if (RES == 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) ptx::ldg256((const float*)ep.res1 + off + 8 * j, rs.f + 8 * j);
  }
  if (RES == 2 || RES == 3) {
#pragma unroll
    for (int j = 0; j < 2; ++j) ptx::ldg256((const h16*)ep.res1 + off + 16 * j, rs.a + 8 * j);
  }
  if (RES == 3) {
#pragma unroll
    for (int j = 0; j < 2; ++j) ptx::ldg256((const h16*)ep.res2 + off + 16 * j, rs.b + 8 * j);
  }
}
__device__ __forceinline__ void add_h16x4(float (&v)[4], uint32_t lo, uint32_t hi, int fp16) {
  const float2 a = unpack_h16x2(lo, fp16), b = unpack_h16x2(hi, fp16);
  v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y;
}
// OUTK: 0 16-bit, 1 fp32, 2 both, 4 16-bit + statistics.  `next()` runs after the last use of rs and before the stores.
template <int ACT, int RES, int OUTK, bool FP16, class Next>
__device__ __forceinline__ void epi_direct_chunk(const uint32_t (&r)[32], uint32_t bias_s, ResD<RES>& rs, float* out32, h16* out16,
                                                 bool row_ok, float& st_s, float& st_q, Next&& next) {
  constexpr int fp16 = FP16 ? 1 : 0;
  constexpr bool W16 = OUTK == 0 || OUTK == 2 || OUTK == 4, W32 = OUTK == 1 || OUTK == 2;
  uint32_t p[W16 ? 16 : 1];
  float4 w[W32 ? 8 : 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias_s) b = lds128(bias_s + 16u * j);                                   // the warp's bias slice, cached in its staging area (broadcast read)
    float v[4] = {__uint_as_float(r[4 * j + 0]) + b.x, __uint_as_float(r[4 * j + 1]) + b.y, __uint_as_float(r[4 * j + 2]) + b.z,
                  __uint_as_float(r[4 * j + 3]) + b.w};
    if (ACT == NPVP_ACT_RELU && RES == 0 && OUTK == 0) {
      // ReLU on the packed 16-bit pairs: rounding is monotonic and 0 is exact, so max(round(x), 0) == round(max(x, 0))
      p[2 * j] = pack_h16x2(v[0], v[1], fp16);
      p[2 * j + 1] = pack_h16x2(v[2], v[3], fp16);
      if (FP16) {
        const __half2 z = __float2half2_rn(0.f);
        *reinterpret_cast<__half2*>(&p[2 * j]) = __hmax2(*reinterpret_cast<__half2*>(&p[2 * j]), z);
        *reinterpret_cast<__half2*>(&p[2 * j + 1]) = __hmax2(*reinterpret_cast<__half2*>(&p[2 * j + 1]), z);
      } else {
        const __nv_bfloat162 z = __float2bfloat162_rn(0.f);
        *reinterpret_cast<__nv_bfloat162*>(&p[2 * j]) = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&p[2 * j]), z);
        *reinterpret_cast<__nv_bfloat162*>(&p[2 * j + 1]) = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&p[2 * j + 1]), z);
      }
      continue;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = act_ct<ACT>(v[e], 0);
    if (RES == 1) {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] += __uint_as_float(rs.f[4 * j + e]);
    }
    if (RES == 2) add_h16x4(v, rs.a[2 * j], rs.a[2 * j + 1], fp16);
    if (RES == 3) {                              // (res1 + res2) first, like the staged epilogue: bit-identical results
      float u[4] = {0.f, 0.f, 0.f, 0.f};
      add_h16x4(u, rs.a[2 * j], rs.a[2 * j + 1], fp16);
      add_h16x4(u, rs.b[2 * j], rs.b[2 * j + 1], fp16);
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] += u[e];
    }
    if (OUTK == 4) {
      st_s += (v[0] + v[1]) + (v[2] + v[3]);
      st_q = fmaf(v[0], v[0], fmaf(v[1], v[1], fmaf(v[2], v[2], fmaf(v[3], v[3], st_q))));
    }
    if (W32) w[j] = make_float4(v[0], v[1], v[2], v[3]);
    if (W16) { p[2 * j] = pack_h16x2(v[0], v[1], fp16); p[2 * j + 1] = pack_h16x2(v[2], v[3], fp16); }
  }
  next();
  if (row_ok) {
    if (W32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) ptx::stg256(out32 + 8 * j, reinterpret_cast<const uint32_t*>(&w[2 * j]));
    }
    if (W16) {
      ptx::stg256(out16, p);
      ptx::stg256(out16 + 16, p + 8);
    }
  }
}
// One tile of one epilogue warp: KCH chunks starting at TMEM address t0 / global column n_first of row m.
//   out_off(n0): element offset of (m, n0) in the outputs;  release(): this warp's part of the accumulator is in registers;
//   next_tile(): residuals of the first chunk of the warp's NEXT tile into rs (RES > 0).
template <int KCH, int ACT, int RES, int OUTK, class OutOff, class Release, class NextTile>
__device__ __forceinline__ void epi_direct_tile(uint32_t t0, int64_t m, bool row_ok, int64_t n_first, int64_t N, const EpiParams& ep, uint32_t bias_s,
                                                int fp16, ResD<RES>& rs, float& st_s, float& st_q, OutOff&& out_off, Release&& release,
                                                NextTile&& next_tile) {
  uint32_t r0[32], r1[32];
  auto process = [&](const uint32_t (&r)[32], int i) {
    const int64_t n0 = n_first + 32 * i;
    if (n0 >= N) return;                                             // warp-uniform; N % 32 == 0 on this path
    const int64_t off = out_off(n0);
    float* o32 = ep.out_f32 ? ep.out_f32 + off : nullptr;
    h16* o16 = ep.out_bf16 ? ep.out_bf16 + off : nullptr;
    auto next = [&]() {
      if (RES > 0) {
        if (i + 1 < KCH && n0 + 32 < N) load_resd<RES>(rs, ep, m, n0 + 32, row_ok);
        else next_tile();
      }
    };
    const uint32_t bs = bias_s ? bias_s + 128u * (uint32_t)i : 0u;
    if (fp16) epi_direct_chunk<ACT, RES, OUTK, true>(r, bs, rs, o32, o16, row_ok, st_s, st_q, next);
    else      epi_direct_chunk<ACT, RES, OUTK, false>(r, bs, rs, o32, o16, row_ok, st_s, st_q, next);
  };
  ptx::tmem_ld_32x32(t0, r0);
#pragma unroll 1
  for (int i = 0; i < KCH; i += 2) {
    ptx::tmem_ld_wait_regs(r0);
    if (i + 1 < KCH) ptx::tmem_ld_32x32(t0 + 32 * (i + 1), r1);
    else release();
    process(r0, i);
    if (i + 1 < KCH) {
      ptx::tmem_ld_wait_regs(r1);
      if (i + 2 < KCH) ptx::tmem_ld_32x32(t0 + 32 * (i + 2), r0);
      else release();
      process(r1, i + 1);
    }
  }
}

template <int BN, int ACT, int RES, int OUT, int CONV>
__global__ void __launch_bounds__(CONV == 1 ? kGemm2Threads + 64 : kGemm2Threads, 1)
gemm_tcgen05_v2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       int64_t M, int64_t N, int64_t K, EpiParams ep, ConvGather cg, TmaConv tc) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr bool DIRECT = OUT >= kDirect;
  constexpr int OUTK = DIRECT ? OUT - kDirect : OUT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  float* staging = (float*)(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* bars = (uint64_t*)(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full_bar = bars + 2 * Cfg::kStages;        // [2]
  uint64_t* tmem_empty_bar = bars + 2 * Cfg::kStages + 2;   // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * Cfg::kStages + 4);
  uint64_t* wres_bar = bars + 2 * Cfg::kStages + 5;         // CONV >= 2, resident weights
  const bool wres = CONV >= 2 && tc.wres;
  const int nst = wres ? tc.nst : Cfg::kStages;             // A-only stages behind the resident weights
  uint8_t* const a_ring = wres ? smem + tc.wres_bytes : smem_a;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_k_blocks = (int)((K + kBK - 1) / kBK);
  const int n_tiles = (int)((N + BN - 1) / BN);
  const int64_t m_tiles = (M + kBM - 1) / kBM;
  const int64_t num_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    if (CONV >= 2) ptx::mbar_init(wres_bar, 1);
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(&full_bar[s], CONV == 1 ? 1 + kGatherThreads : 1);   // TMA thread (+ every gather thread in gather mode)
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], kEpiWarps);      // one arrival per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t empty0 = ptx::smem_u32(empty_bar);
      bool freed = false;                                            // empty_bar[stage] already seen complete (tested one k-block ahead)
      if (wres) {                                                    // the live weight blocks, once per CTA
        ptx::mbar_arrive_expect_tx(wres_bar, (uint32_t)tc.wres_bytes);
        uint32_t off = 0;
        for (int tp = 0; tp < tc.taps; ++tp) {
          int lo = 0, hi = (int)N;
          if (CONV == 3) convt_cols(tp, tc.cout, 0, (int)N, lo, hi);
          for (int kc = 0; kc < tc.kb_per_tap; ++kc)
            for (int r0 = lo; r0 < hi; r0 += tc.wbox) {
              ptx::tma_load_2d(smem + off, &tmap_b, wres_bar, (tp * tc.kb_per_tap + kc) * kBK, r0);
              off += (uint32_t)tc.wbox * 128u;
            }
        }
      }
      // Dense A streams from HBM exactly once, so every first touch of an A box is a DRAM-latency load: with 4 stages (192 KB)
      // in flight per SM the ring sustained ~60 B/clk against the 96 B/clk a 128 x 256 tile needs (r02 timeline: operands
      // landed 2000-3000 clk after their stage was freed).  An L2 prefetch kPfDist k-blocks ahead of the ring turns those into
      // L2 hits without spending shared memory.
      constexpr int kPfDist = 8;
      int64_t pf_t = blockIdx.x;
      int pf_kb = 0;
      auto pf_advance = [&]() { if (++pf_kb == num_k_blocks) { pf_kb = 0; pf_t += gridDim.x; } };
      if (!CONV && tc.pf) for (int i = 0; i < kPfDist; ++i) pf_advance();
      for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_blk = (int)(t / n_tiles), n_blk = (int)(t % n_tiles);
        if (CONV >= 2) {
          // shifted-window A tiles: tile = 128 consecutive pixels starting at row y0 of frame f0 (or whole frames from f0)
          const int64_t m0 = (int64_t)m_blk * kBM;
          const int f0 = (int)(m0 / tc.HW);
          const int y0 = (int)(m0 - (int64_t)f0 * tc.HW) / tc.W;
          for (int tp = 0; tp < tc.taps; ++tp) {
            if (CONV == 3) {
              int lo, hi;
              convt_cols(tp, tc.cout, n_blk * BN, BN, lo, hi);
              if (hi <= lo) continue;                                // this tap feeds none of the tile's output phases
            }
            const int ty = tp / tc.KW, tx = tp - ty * tc.KW;
            for (int kc = 0; kc < tc.kb_per_tap; ++kc) {
              if (!freed) ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
              NPVP_TRACE(0);
              int ns = stage + 1;
              uint32_t np = phase;
              if (ns == nst) { ns = 0; np ^= 1; }
              freed = ptx::mbar_test(empty0 + 8u * (uint32_t)ns, np ^ 1);
              ptx::mbar_arrive_expect_tx(&full_bar[stage], wres ? Cfg::kABytes : Cfg::kStageBytes);
              ptx::tma_load_4d(a_ring + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kc * kBK, tx - tc.pad, y0 + ty - tc.pad, f0);
              if (!wres) ptx::tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], (tp * tc.kb_per_tap + kc) * kBK, n_blk * BN);
              stage = ns; phase = np;
            }
          }
          continue;
        }
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          if (!freed) ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          NPVP_TRACE(0);
          int ns = stage + 1;
          uint32_t np = phase;
          if (ns == Cfg::kStages) { ns = 0; np ^= 1; }
          freed = ptx::mbar_test(empty0 + 8u * (uint32_t)ns, np ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], CONV ? Cfg::kBBytes : Cfg::kStageBytes);
          if (!CONV) ptx::tma_load_2d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * kBK, m_blk * kBM);
          ptx::tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * kBK, n_blk * BN);
          if (!CONV && tc.pf) {
            if (pf_t < num_tiles) ptx::tma_prefetch_2d(&tmap_a, pf_kb * kBK, (int)(pf_t / n_tiles) * kBM);
            pf_advance();
          }
          stage = ns; phase = np;
        }
      }
    }
  } else if (CONV == 1 && (warp == 2 || warp == 3 || warp >= 12)) {
    // ---------------- implicit-GEMM A gather (128 threads: warps 2, 3, 12, 13) ----------------
    // Lane layout: 8 lanes cover the 128 bytes of one A row (one 16-byte chunk each), 16 rows per pass, 8 passes per
    // k-block.  The producer is bound by the latency of its own instruction stream, not by memory: the first version
    // (64 threads, two whole rows each, full address arithmetic per row and k-block) executed ~470 dependent instructions
    // per k-block = 2150 clk against 540 clk of MMA work (r01 ncu of the decoder layers, tensor pipe 26% active, MMA warp
    // spinning on the full barrier).  So the per-row arithmetic is hoisted into a table - whenever the k-loop enters a new
    // filter tap each thread computes the source pixel of ONE row for that tap into shared memory - and the per-k-block
    // work is 8 table reads + 8 cp.async per thread.
    const int pt = warp < 4 ? (int)threadIdx.x - 64 : (int)threadIdx.x - kGemm2Threads + 64;
    const int chunk = pt & 7, rslot = pt >> 3;                       // rslot: row within a 16-row pass
    const int C = cg.C;
    const int tpk = C < kBK ? kBK / C : 1;                           // taps per k-block (2 when C == 32)
    const int kb_per_tap = C > kBK ? C / kBK : 1;                    // k-blocks per tap (C > 64)
    const int taps = cg.KH * cg.KW;
    const uint32_t hw_out = (uint32_t)(cg.Ho * cg.Wo);
    uint32_t* table = reinterpret_cast<uint32_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kStagingBytes + 256);   // [2][tpk][128]
    const int my_tt = C < kBK ? (chunk * 8) / C : 0;                 // which of the k-block's taps this lane's chunk belongs to
    const int my_c = C < kBK ? (chunk * 8) % C : chunk * 8;          // channel offset of the chunk inside that tap
    const uint32_t dst_off = (uint32_t)(rslot * 128 + ((chunk ^ (rslot & 7)) << 4));   // row & 7 == rslot & 7 for every pass
    constexpr uint32_t kNoPixel = 0xffffffffu;
    const int H = cg.H, W = cg.W, H2 = cg.H >> 1, W2 = cg.W >> 1;
    int stage = 0, trail = 0, inflight = 0, buf = 0;
    uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = (int)(t / n_tiles);
      uint32_t fbase;                                                 // first pixel of the row's frame (all pixel indices fit 32 bits)
      int iy0, ix0;
      bool rok;
      {
        const int64_t m = (int64_t)m_blk * kBM + pt;                  // this thread's table row
        rok = m < M;
        const uint32_t mm = rok ? (uint32_t)m : 0u;
        const uint32_t f = mm / hw_out, rem = mm - f * hw_out;
        const uint32_t oy = rem / (uint32_t)cg.Wo;
        fbase = f * (uint32_t)(H * W);
        iy0 = (int)oy * cg.stride - cg.pad;
        ix0 = (int)(rem - oy * (uint32_t)cg.Wo) * cg.stride - cg.pad;
      }
      int tap_next = 0, ky = 0, kx = 0, kb_left = 0, c_base = 0;     // (ky, kx) of tap_next
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        if (kb_left == 0) {                                           // the k-loop enters new tap(s): rebuild the pixel table
          buf ^= 1;
          for (int tt = 0; tt < tpk; ++tt) {
            {
              int iy = iy0 + ky, ix = ix0 + kx;
              bool valid = rok && tap_next < taps;
              if (cg.pad_mode == NPVP_PAD_REFLECT) { iy = conv_reflect(iy, H); ix = conv_reflect(ix, W); }
              else if (cg.pad_mode == NPVP_PAD_REPLICATE) { iy = min(max(iy, 0), H - 1); ix = min(max(ix, 0), W - 1); }
              else valid = valid && (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
              const uint32_t pix = cg.phase_major
                  ? fbase + (uint32_t)((((iy >> 1) * W2 + (ix >> 1)) << 2) + ((iy & 1) << 1) + (ix & 1))
                  : fbase + (uint32_t)(iy * W + ix);
              table[(buf * tpk + tt) * kBM + pt] = valid ? pix : kNoPixel;
            }
            ++tap_next;
            if (++kx == cg.KW) { kx = 0; ++ky; }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");             // the four gather warps only
          kb_left = kb_per_tap;
          c_base = 0;
        }
        // all 8 source pixels first: the cp.async asm statements below are compiler barriers for shared-memory loads
        const uint32_t* tb = table + (buf * tpk + my_tt) * kBM + rslot;
        uint32_t pix[kBM / kGatherRowsPerPass];
#pragma unroll
        for (int p = 0; p < kBM / kGatherRowsPerPass; ++p) pix[p] = tb[p * kGatherRowsPerPass];
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        const uint32_t a_dst = ptx::smem_u32(smem_a + stage * Cfg::kABytes) + dst_off;
        const h16* xb = cg.x + c_base + my_c;
#pragma unroll
        for (int p = 0; p < kBM / kGatherRowsPerPass; ++p) {
          const bool valid = pix[p] != kNoPixel;
          cp_async_16(a_dst + (uint32_t)p * (kGatherRowsPerPass * 128u), xb + (size_t)(valid ? pix[p] : 0u) * (uint32_t)C, valid ? 16u : 0u);
        }
        cp_async_commit();
        --kb_left;
        c_base += kBK;
        ++inflight;
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        if (inflight > kGatherLag) {                                  // oldest group has landed: publish its stage
          cp_async_wait<kGatherLag>();
          ptx::fence_proxy_async();
          mbar_arrive(&full_bar[trail]);
          if (++trail == Cfg::kStages) trail = 0;
          --inflight;
        }
      }
    }
    cp_async_wait<0>();
    ptx::fence_proxy_async();
    while (inflight > 0) {
      mbar_arrive(&full_bar[trail]);
      if (++trail == Cfg::kStages) trail = 0;
      --inflight;
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (ptx::elect_one()) {
      const uint32_t idesc = make_idesc_f16kind(kBM, BN, ep.fp16);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      // stage s: operands at a_lo0 + s * kAStep / b_lo0 + s * kBStep (descriptor low words), full / empty barriers 8 s bytes on
      const uint32_t a_lo0 = smem_desc_lo(ptx::smem_u32(a_ring)), b_lo0 = smem_desc_lo(ptx::smem_u32(smem_b));
      const uint32_t full0 = ptx::smem_u32(full_bar);
      constexpr uint32_t kAStep = Cfg::kABytes >> 4, kBStep = Cfg::kBBytes >> 4;
      bool landed = false;                                           // full_bar[stage] already seen complete (tested one k-block ahead)
      if (wres) {
        ptx::mbar_wait(wres_bar, 0);
        ptx::tc_fence_after();
      }
      for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);       // epilogue has drained this accumulator
        NPVP_TRACE(3);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        if (CONV >= 2) {
          const int n_blk = (int)(t % n_tiles);
          uint32_t started = 0;                                      // tap 0 covers every column of the tile: it initialises the accumulator
          uint32_t woff = 0;                                         // resident weights: blocks lie in (tap, k-block) order, live rows only
          for (int tp = 0; tp < tc.taps; ++tp) {
            int lo = 0, hi = wres ? (int)N : BN;
            if (CONV == 3) {
              convt_cols(tp, tc.cout, n_blk * BN, wres ? (int)N : BN, lo, hi);
              if (hi <= lo) continue;
            }
            const uint32_t idesc_t = (CONV == 3 || wres) ? make_idesc_f16kind(kBM, hi - lo, ep.fp16) : idesc;
            for (int kc = 0; kc < tc.kb_per_tap; ++kc) {
              if (!landed) ptx::mbar_wait(&full_bar[stage], phase);
              NPVP_TRACE(1);
              ptx::tc_fence_after();
              int ns = stage + 1;
              uint32_t np = phase;
              if (ns == nst) { ns = 0; np ^= 1; }
              landed = ptx::mbar_test(full0 + 8u * (uint32_t)ns, np);   // the next stage's barrier: its round trip hides behind the MMA issue
              const uint32_t b_lo = wres ? smem_desc_lo(ptx::smem_u32(smem) + woff) : b_lo0 + (uint32_t)stage * kBStep + (uint32_t)lo * 8u;   // lo % 8 == 0: whole swizzle atoms
              ptx::umma_f16_x4(tmem_d + (uint32_t)lo, a_lo0 + (uint32_t)stage * kAStep, b_lo, idesc_t, started);
              started = 1;
              woff += (uint32_t)(hi - lo) * 128u;
              ptx::umma_commit(&empty_bar[stage]);
              NPVP_TRACE(2);
              stage = ns; phase = np;
            }
          }
        } else {
          for (int kb = 0; kb < num_k_blocks; ++kb) {
            if (!landed) ptx::mbar_wait(&full_bar[stage], phase);
            NPVP_TRACE(1);
            ptx::tc_fence_after();
            int ns = stage + 1;
            uint32_t np = phase;
            if (ns == Cfg::kStages) { ns = 0; np ^= 1; }
            landed = ptx::mbar_test(full0 + 8u * (uint32_t)ns, np);     // the next stage's barrier: its round trip hides behind the MMA issue
            ptx::umma_f16_x4(tmem_d, a_lo0 + (uint32_t)stage * kAStep, b_lo0 + (uint32_t)stage * kBStep, idesc, (uint32_t)kb);
            ptx::umma_commit(&empty_bar[stage]);
            NPVP_TRACE(2);
            stage = ns; phase = np;
          }
        }
        ptx::umma_commit(&tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 4 + kEpiWarps) {
    // ---------------- epilogue ----------------
    const int quad = warp & 3;                                      // TMEM lane quadrant of this warp
    const int half = (warp - 4) >> 2;                               // which half of the tile's columns
    const uint32_t stg_addr = ptx::smem_u32(staging + (warp - 4) * (32 * 32));
    int acc = 0;
    uint32_t acc_phase = 0;
    const int sub_row = lane >> 3, chunk = lane & 7;                // coalesced layout: 4 rows x 8 float4 per instruction
    const bool has_bias = ep.bias != nullptr;
    const int fp16 = ep.out_fp16;                                  // 16-bit type of the OUTPUT (and of 16-bit residuals)
    const float alpha = ep.alpha;
    const float relu_floor = ep.post_relu ? 0.f : -INFINITY;
    constexpr int kColsPerWarp = BN / 2;
    ResRaw<RES> rr;
    ResD<RES> rd;                                                    // direct epilogue: residuals of the chunk about to be processed
    int bias_blk = -1;                                               // n-block whose bias slice the staging area holds (direct epilogue)
    int64_t rd_tile = -1;                                            // tile whose first chunk rd was loaded for
    int64_t rr_tile = -1;                                            // which (tile, tile column) rr holds
    int rr_c = -1;
    auto res_prefetch = [&](int64_t tt, int c) {                    // residuals of the chunk at tile column c of tile tt
      rr_tile = tt; rr_c = c;
      if (tt >= num_tiles) return;
      const int64_t nn = (tt % n_tiles) * BN + c + chunk * 4;
      load_res<RES>(rr, ep, sub_row, (tt / n_tiles) * kBM + quad * 32, M, nn < N, nn);
    };
    for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = (int)(t / n_tiles), n_blk = (int)(t % n_tiles);
      // bias of the first column chunk: fetched BEFORE waiting for the accumulator, later chunks one iteration ahead, so the
      // L2 latency never sits on the epilogue's critical path (r01 profile: long-scoreboard stalls on the per-chunk bias load)
      auto load_bias = [&](int c) {
        const int64_t nb = (int64_t)n_blk * BN + c + chunk * 4;
        return (has_bias && nb < N) ? __ldg(reinterpret_cast<const float4*>(ep.bias + nb)) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      if (DIRECT) {
        // (CONV == 3, transposed conv: row m = input pixel, a chunk = 32 channels of ONE output pixel (2y+py, 2x+px) since
        //  Cout % 32 == 0; column n0 = (phase q, co) with q enumerating (py,px) = (0,0), (0,1), (1,1), (1,0))
        constexpr int kChunks = kColsPerWarp / 32;
        const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * kColsPerWarp);
        const int64_t m = (int64_t)m_blk * kBM + quad * 32 + lane;
        const int64_t nw = (int64_t)n_blk * BN + half * kColsPerWarp;
        auto out_off = [&](int64_t n0) -> int64_t {
          if (CONV == 3) {
            const uint32_t cout = (uint32_t)tc.cout, mm = (uint32_t)m, nn = (uint32_t)n0;
            const uint32_t q = (nn >= cout) + (nn >= 2u * cout) + (nn >= 3u * cout), co = nn - q * cout;
            return (int64_t)cout * (int64_t)(4u * mm - 2u * (mm & ((uint32_t)tc.W - 1u))) +
                   (int64_t)((q >> 1) * (2u * (uint32_t)tc.W * cout) + (((q + 1) >> 1) & 1u) * cout + co);
          }
          return m * ep.ld_out + n0;
        };
        auto release = [&]() {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        };
        auto next_tile = [&]() {
          const int64_t tn = t + gridDim.x;
          rd_tile = tn;
          if (tn < num_tiles) {
            const int64_t mn = (tn / n_tiles) * kBM + quad * 32 + lane;
            load_resd<RES>(rd, ep, mn, (tn % n_tiles) * BN + half * kColsPerWarp, mn < M);
          }
        };
        if (RES > 0 && rd_tile != t) load_resd<RES>(rd, ep, m, nw, m < M && nw < N);   // first tile of this warp
        float st_s = 0.f, st_q = 0.f;
        // The warp's bias slice (kColsPerWarp floats) lives in its (otherwise unused) staging area and is refreshed only when the
        // tile's n-block changes: the direct layout needs 32 bias values per thread and chunk, and eight warp-uniform LDG.128 per
        // chunk were the top stall of the store-bound epilogues (r02 ncu of the last transposed conv: 22 % of the samples on the
        // FADDs that consume them).
        if (has_bias && n_blk != bias_blk) {
          __syncwarp();
          if (lane * 4 < kColsPerWarp) {
            const int64_t nb = nw + lane * 4;
            const float4 b = nb < N ? __ldg(reinterpret_cast<const float4*>(ep.bias + nb)) : make_float4(0.f, 0.f, 0.f, 0.f);
            sts128(stg_addr + 16u * (uint32_t)lane, __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
          }
          __syncwarp();
          bias_blk = n_blk;
        }
        ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
        if (threadIdx.x == 128) NPVP_TRACE(4);
        ptx::tc_fence_after();
        epi_direct_tile<kChunks, ACT, RES, OUTK>(t0, m, m < M, nw, N, ep, has_bias ? stg_addr : 0u, fp16, rd, st_s, st_q, out_off, release, next_tile);
        if (threadIdx.x == 128) NPVP_TRACE(5);
        if (OUTK == kOutStats16) {
          // this warp's 32 rows x 128 columns lie inside one 64-row frame: one (sum, sum of squares) slot per warp and tile
          st_s = warp_sum(st_s);
          st_q = warp_sum(st_q);
          if (lane == 0 && (int64_t)m_blk * kBM + quad * 32 < M) {
            const int64_t frame = (int64_t)m_blk * 2 + (quad >> 1);
            const int slot = ((n_blk * 2 + half) << 1) + (quad & 1);
            float2* dst = reinterpret_cast<float2*>(ep.frame_stats) + frame * (n_tiles * 4) + slot;
            *dst = make_float2(st_s, st_q);
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      float4 b_next = load_bias(half * kColsPerWarp);
      if (RES > 0 && (rr_tile != t || rr_c != half * kColsPerWarp)) res_prefetch(t, half * kColsPerWarp);   // (normally done by the previous tile's last chunk)
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      if (threadIdx.x == 128) NPVP_TRACE(4);
      ptx::tc_fence_after();
      const int64_t m_base = (int64_t)m_blk * kBM + quad * 32;
      float st_s = 0.f, st_q = 0.f;                                  // OUT == kOutStats16 only
#pragma unroll 1
      for (int c = half * kColsPerWarp; c < (half + 1) * kColsPerWarp; c += 32) {
        const float4 b4 = b_next;
        if (c + 32 < (half + 1) * kColsPerWarp) b_next = load_bias(c + 32);
        const int64_t n0 = (int64_t)n_blk * BN + c;
        if (n0 >= N) break;                                          // warp-uniform
        if (RES > 0 && (rr_tile != t || rr_c != c)) res_prefetch(t, c);   // not prefetched (a tile whose first chunk lay past N came in between)
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + c), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)                                  // float4 chunk j of row `lane` -> slot j ^ (lane & 7)
          sts128(stg_addr + (uint32_t)((lane * 32 + ((j ^ (lane & 7)) << 2)) * 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        const int64_t n = n0 + chunk * 4;
        const bool col_ok = n < N;                                   // N % 4 == 0 on this path
        // (residuals may alias the output - in-place residual-stream update - so a chunk's residuals are loaded before any of
        //  ITS stores; the prefetch of the following chunk reads other columns / another tile)
        auto next = [&]() {
          if (RES > 0) {
            const int cn = c + 32;
            if (cn < (half + 1) * kColsPerWarp && (int64_t)n_blk * BN + cn < N) res_prefetch(t, cn);
            else res_prefetch(t + gridDim.x, half * kColsPerWarp);
          }
        };
        if (fp16) epi_chunk<ACT, RES, OUT, true>(stg_addr, sub_row, chunk, m_base, M, col_ok, n, b4, alpha, relu_floor, ep, st_s, st_q, rr, next);
        else      epi_chunk<ACT, RES, OUT, false>(stg_addr, sub_row, chunk, m_base, M, col_ok, n, b4, alpha, relu_floor, ep, st_s, st_q, rr, next);
        __syncwarp();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (threadIdx.x == 128) NPVP_TRACE(5);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (OUT == kOutStats16) {
        // this warp's 32 rows x 128 columns lie inside one 64-row frame: one (sum, sum of squares) slot per warp and tile
        st_s = warp_sum(st_s);
        st_q = warp_sum(st_q);
        if (lane == 0 && m_base < M) {                               // (the last tile may hold a single frame)
          const int64_t frame = (int64_t)m_blk * 2 + (quad >> 1);
          const int slot = ((n_blk * 2 + half) << 1) + (quad & 1);
          float2* dst = reinterpret_cast<float2*>(ep.frame_stats) + frame * (n_tiles * 4) + slot;
          *dst = make_float2(st_s, st_q);
        }
      }
    }
  }
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// =============================================================================================
// tcgen05 GEMM kernel, 2-CTA variant (cta_group::2): a pair of SMs computes a 256 x 256 tile
// =============================================================================================
// r01 finding: the 1-CTA kernel is bound by shared-memory bandwidth (per 64-wide k-block a 128x256 tile makes TMA
// write 48 KB and the tensor core read 48 KB, the epilogue transpose adds ~33 KB: 129 KB against 128 B/clk x 512 clk).
// With cta_group::2 the two CTAs of a cluster each stage their own 128 rows of A and HALF of the B tile (128 of the 256
// N-rows); one tcgen05.mma issued by the leader drives both tensor cores (UMMA M = 256) and B is shared between the
// pair in hardware, so operand traffic per SM drops from 96 KB to 64 KB per k-block.
//   - both CTAs run a TMA producer; all complete_tx land on the LEADER's full barrier (peer bit cleared in the address)
//   - only the leader issues MMAs; tcgen05.commit ... multicast::cluster releases the smem stage / publishes the
//     accumulator in both CTAs
//   - each CTA's 8 epilogue warps drain their own TMEM half and arrive (remotely for the follower) on the leader's
//     tmem_empty barrier
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

struct Gemm3Cfg {
  static constexpr int BN = 256;
  static constexpr int kABytes = kBM * kBK * 2;            // this CTA's 128 rows of A
  static constexpr int kBBytes = (BN / 2) * kBK * 2;       // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;    // 32 KB
  static constexpr int kStages = 6;
  static constexpr int kStagingBytes = kEpiWarps * 32 * 32 * 4;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 + 256;
  static constexpr uint32_t kTmemCols = 512;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(map), "r"(ptx::smem_u32(bar) & kPeerBitMask), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_x4_2cta(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {   // see ptx::umma_f16_x4
  asm volatile(
      "{\n\t.reg .pred p, pt;\n\t.reg .b64 da, db;\n\t.reg .b32 la, lb;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 pt, %5, %5;\n\t"
      "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t"
      "add.u32 la, %1, 2;\n\tadd.u32 lb, %2, 2;\n\tmov.b64 da, {la, %5};\n\tmov.b64 db, {lb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.u32 la, %1, 4;\n\tadd.u32 lb, %2, 4;\n\tmov.b64 da, {la, %5};\n\tmov.b64 db, {lb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.u32 la, %1, 6;\n\tadd.u32 lb, %2, 6;\n\tmov.b64 da, {la, %5};\n\tmov.b64 db, {lb, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(0x40004040u)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {     // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(ptx::smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

template <int ACT, int RES, int OUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemm2Threads, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         int64_t M, int64_t N, int64_t K, EpiParams ep) {
  using Cfg = Gemm3Cfg;
  constexpr int BN = Cfg::BN;
  constexpr bool DIRECT = OUT >= kDirect;
  constexpr int OUTK = DIRECT ? OUT - kDirect : OUT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  float* staging = (float*)(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* bars = (uint64_t*)(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full_bar = bars + 2 * Cfg::kStages;        // [2]
  uint64_t* tmem_empty_bar = bars + 2 * Cfg::kStages + 2;   // [2]  (the leader's copy is the live one)
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * Cfg::kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_k_blocks = (int)((K + kBK - 1) / kBK);
  const int n_tiles = (int)((N + BN - 1) / BN);
  const int64_t m_tiles = (M + 2 * kBM - 1) / (2 * kBM);
  const int64_t num_tiles = m_tiles * n_tiles;
  const int64_t pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], 2 * kEpiWarps);   // epilogue warps of both CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_slot)), "r"(Cfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  ptx::tc_fence_before();
  __syncthreads();                                           // CTA-local: the TMEM address written by tcgen05.alloc is visible (racecheck-clean)
  cluster_sync_all();                                        // barriers of both CTAs are initialised before any remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs) ----------------
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t empty0 = ptx::smem_u32(empty_bar);
      bool freed = false;                                            // empty_bar[stage] already seen complete (tested one k-block ahead)
      for (int64_t t = pair; t < num_tiles; t += num_pairs) {
        const int m_blk = (int)(t / n_tiles), n_blk = (int)(t % n_tiles);
        const int m_row = m_blk * 2 * kBM + (int)rank * kBM;
        const int n_row = n_blk * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          if (!freed) ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          NPVP_TRACE(0);
          int ns = stage + 1;
          uint32_t np = phase;
          if (ns == Cfg::kStages) { ns = 0; np ^= 1; }
          freed = ptx::mbar_test(empty0 + 8u * (uint32_t)ns, np ^ 1);
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          tma_load_2d_2sm(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * kBK, m_row);
          tma_load_2d_2sm(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * kBK, n_row);
          stage = ns; phase = np;
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (leader CTA only) ----------------
    if (leader && ptx::elect_one()) {
      const uint32_t idesc = make_idesc_f16kind(2 * kBM, BN, ep.fp16);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t a_lo0 = smem_desc_lo(ptx::smem_u32(smem_a)), b_lo0 = smem_desc_lo(ptx::smem_u32(smem_b));
      const uint32_t full0 = ptx::smem_u32(full_bar);
      constexpr uint32_t kAStep = Cfg::kABytes >> 4, kBStep = Cfg::kBBytes >> 4;
      bool landed = false;                                           // full_bar[stage] already seen complete (tested one k-block ahead)
      for (int64_t t = pair; t < num_tiles; t += num_pairs) {
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        NPVP_TRACE(3);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          if (!landed) ptx::mbar_wait(&full_bar[stage], phase);
          NPVP_TRACE(1);
          ptx::tc_fence_after();
          int ns = stage + 1;
          uint32_t np = phase;
          if (ns == Cfg::kStages) { ns = 0; np ^= 1; }
          landed = ptx::mbar_test(full0 + 8u * (uint32_t)ns, np);
          umma_f16_x4_2cta(tmem_d, a_lo0 + (uint32_t)stage * kAStep, b_lo0 + (uint32_t)stage * kBStep, idesc, (uint32_t)kb);
          umma_commit_2cta(&empty_bar[stage]);
          NPVP_TRACE(2);
          stage = ns; phase = np;
        }
        umma_commit_2cta(&tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue (both CTAs, own 128 rows) ----------------
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t stg_addr = ptx::smem_u32(staging + (warp - 4) * (32 * 32));
    int acc = 0;
    uint32_t acc_phase = 0;
    const int sub_row = lane >> 3, chunk = lane & 7;
    const bool has_bias = ep.bias != nullptr;
    const int fp16 = ep.out_fp16;                                  // 16-bit type of the OUTPUT (and of 16-bit residuals)
    const float alpha = ep.alpha;
    const float relu_floor = ep.post_relu ? 0.f : -INFINITY;
    constexpr int kColsPerWarp = BN / 2;
    ResRaw<RES> rr;
    ResD<RES> rd;                                                    // direct epilogue: residuals of the chunk about to be processed
    int bias_blk = -1;                                               // n-block whose bias slice the staging area holds (direct epilogue)
    int64_t rd_tile = -1;                                            // tile whose first chunk rd was loaded for
    int64_t rr_tile = -1;
    int rr_c = -1;
    auto res_prefetch = [&](int64_t tt, int c) {
      rr_tile = tt; rr_c = c;
      if (tt >= num_tiles) return;
      const int64_t nn = (tt % n_tiles) * BN + c + chunk * 4;
      load_res<RES>(rr, ep, sub_row, (tt / n_tiles) * 2 * kBM + (int64_t)rank * kBM + quad * 32, M, nn < N, nn);
    };
    const uint32_t empty_remote[2] = {mapa_u32(ptx::smem_u32(&tmem_empty_bar[0]), 0), mapa_u32(ptx::smem_u32(&tmem_empty_bar[1]), 0)};
    for (int64_t t = pair; t < num_tiles; t += num_pairs) {
      const int m_blk = (int)(t / n_tiles), n_blk = (int)(t % n_tiles);
      auto load_bias = [&](int c) {
        const int64_t nb = (int64_t)n_blk * BN + c + chunk * 4;
        return (has_bias && nb < N) ? __ldg(reinterpret_cast<const float4*>(ep.bias + nb)) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      if (DIRECT) {
        constexpr int kChunks = kColsPerWarp / 32;
        const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * kColsPerWarp);
        const int64_t m = (int64_t)m_blk * 2 * kBM + (int64_t)rank * kBM + quad * 32 + lane;
        const int64_t nw = (int64_t)n_blk * BN + half * kColsPerWarp;
        auto out_off = [&](int64_t n0) -> int64_t { return m * ep.ld_out + n0; };
        auto release = [&]() {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(empty_remote[acc]);     // the leader's barrier counts both CTAs' epilogue warps
        };
        auto next_tile = [&]() {
          const int64_t tn = t + num_pairs;
          rd_tile = tn;
          if (tn < num_tiles) {
            const int64_t mn = (tn / n_tiles) * 2 * kBM + (int64_t)rank * kBM + quad * 32 + lane;
            load_resd<RES>(rd, ep, mn, (tn % n_tiles) * BN + half * kColsPerWarp, mn < M);
          }
        };
        if (RES > 0 && rd_tile != t) load_resd<RES>(rd, ep, m, nw, m < M && nw < N);
        float st_s = 0.f, st_q = 0.f;
        // The warp's bias slice (kColsPerWarp floats) lives in its (otherwise unused) staging area and is refreshed only when the
        // tile's n-block changes: the direct layout needs 32 bias values per thread and chunk, and eight warp-uniform LDG.128 per
        // chunk were the top stall of the store-bound epilogues (r02 ncu of the last transposed conv: 22 % of the samples on the
        // FADDs that consume them).
        if (has_bias && n_blk != bias_blk) {
          __syncwarp();
          if (lane * 4 < kColsPerWarp) {
            const int64_t nb = nw + lane * 4;
            const float4 b = nb < N ? __ldg(reinterpret_cast<const float4*>(ep.bias + nb)) : make_float4(0.f, 0.f, 0.f, 0.f);
            sts128(stg_addr + 16u * (uint32_t)lane, __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
          }
          __syncwarp();
          bias_blk = n_blk;
        }
        ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
        if (threadIdx.x == 128) NPVP_TRACE(4);
        ptx::tc_fence_after();
        epi_direct_tile<kChunks, ACT, RES, OUTK>(t0, m, m < M, nw, N, ep, has_bias ? stg_addr : 0u, fp16, rd, st_s, st_q, out_off, release, next_tile);
        if (threadIdx.x == 128) NPVP_TRACE(5);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      float4 b_next = load_bias(half * kColsPerWarp);
      if (RES > 0 && (rr_tile != t || rr_c != half * kColsPerWarp)) res_prefetch(t, half * kColsPerWarp);
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      if (threadIdx.x == 128) NPVP_TRACE(4);
      ptx::tc_fence_after();
      const int64_t m_base = (int64_t)m_blk * 2 * kBM + (int64_t)rank * kBM + quad * 32;
      float st_s = 0.f, st_q = 0.f;                                  // (frame statistics are a 1-CTA kernel feature)
#pragma unroll 1
      for (int c = half * kColsPerWarp; c < (half + 1) * kColsPerWarp; c += 32) {
        const float4 b4 = b_next;
        if (c + 32 < (half + 1) * kColsPerWarp) b_next = load_bias(c + 32);
        const int64_t n0 = (int64_t)n_blk * BN + c;
        if (n0 >= N) break;
        if (RES > 0 && (rr_tile != t || rr_c != c)) res_prefetch(t, c);
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + c), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(stg_addr + (uint32_t)((lane * 32 + ((j ^ (lane & 7)) << 2)) * 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        const int64_t n = n0 + chunk * 4;
        const bool col_ok = n < N;
        auto next = [&]() {
          if (RES > 0) {
            const int cn = c + 32;
            if (cn < (half + 1) * kColsPerWarp && (int64_t)n_blk * BN + cn < N) res_prefetch(t, cn);
            else res_prefetch(t + num_pairs, half * kColsPerWarp);
          }
        };
        if (fp16) epi_chunk<ACT, RES, OUT, true>(stg_addr, sub_row, chunk, m_base, M, col_ok, n, b4, alpha, relu_floor, ep, st_s, st_q, rr, next);
        else      epi_chunk<ACT, RES, OUT, false>(stg_addr, sub_row, chunk, m_base, M, col_ok, n, b4, alpha, relu_floor, ep, st_s, st_q, rr, next);
        __syncwarp();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(empty_remote[acc]);       // the leader's barrier counts both CTAs' epilogue warps
      if (threadIdx.x == 128) NPVP_TRACE(5);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  ptx::tc_fence_before();
  cluster_sync_all();                                              // peer smem / TMEM stay valid until both CTAs are done
  if (warp == 2) {
    ptx::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
  }
}

// =============================================================================================
// SIMT kernels
// =============================================================================================
// 64x64 tile, 16-deep k-slab, 256 threads, 4x4 outputs per thread.  TA in {bf16, float}.
template <typename TA>
__device__ __forceinline__ float to_f32(TA v, int fp16);
template <> __device__ __forceinline__ float to_f32<float>(float v, int) { return v; }
template <> __device__ __forceinline__ float to_f32<h16>(h16 v, int fp16) { return h16_to_float(v, fp16); }

template <typename TA>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, int64_t lda, const TA* __restrict__ W, int64_t ldw, int64_t M, int64_t N,
                 int64_t K, EpiParams ep) {
  __shared__ float As[16][64 + 1];
  __shared__ float Ws[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * 64, n0 = (int64_t)blockIdx.x * 64;
  float acc[4][4] = {};
  for (int64_t k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, c = i & 15;
      const int64_t m = m0 + r, n = n0 + r, k = k0 + c;
      As[c][r] = (m < M && k < K) ? to_f32<TA>(A[m * lda + k], ep.fp16) : 0.0f;
      Ws[c][r] = (n < N && k < K) ? to_f32<TA>(W[n * ldw + k], ep.fp16) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const float v = epi_value(ep, acc[i][j], m, n);
      if (ep.out_f32) ep.out_f32[m * ep.ld_out + n] = v;
      if (ep.out_bf16) ep.out_bf16[m * ep.ld_out + n] = float_to_h16(v, ep.out_fp16);
    }
  }
}

// =============================================================================================
// host side
// =============================================================================================

// 2-D bf16 tensor map: inner dim = K (contiguous), outer dim = rows; box = 64 x box_rows; 128B swizzle; OOB -> 0.
static int make_tmap_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows, int fp16) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { npvp_set_error("cuTensorMapEncodeTiled entry point unavailable"); return NPVP_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { npvp_set_error("cuTensorMapEncodeTiled failed (%d): rows=%lld K=%lld ld=%lld", (int)r, (long long)rows, (long long)K, (long long)ld); return NPVP_ERR_CUDA; }
  return NPVP_OK;
}

// 4-D 16-bit tensor map over channels-last activations [frames, H, W, C]: box = 64 channels x (bw x bh x bf = 128 pixels),
// 128B swizzle, out-of-range pixels (the zero padding of the convolution, frames past the end) read as zeros.
static int make_tmap_nhwc(CUtensorMap* map, const void* base, int64_t frames, int H, int W, int C, int fp16) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { npvp_set_error("cuTensorMapEncodeTiled entry point unavailable"); return NPVP_ERR_CUDA; }
  const int HW = H * W;
  const int bh = HW >= kBM ? kBM / W : H, bf = HW >= kBM ? 1 : kBM / HW;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)frames};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)HW * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)W, (cuuint32_t)bh, (cuuint32_t)bf};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { npvp_set_error("cuTensorMapEncodeTiled (NHWC) failed (%d): frames=%lld H=%d W=%d C=%d", (int)r, (long long)frames, H, W, C); return NPVP_ERR_CUDA; }
  return NPVP_OK;
}
// geometry the shifted-window (TMA) A path can tile: 128-pixel tiles are whole rows of one frame or whole frames
static bool tma_conv_geometry(int H, int W, int C) {
  const int HW = H * W;
  return C % kBK == 0 && W >= 1 && W <= kBM && kBM % W == 0 && (HW % kBM == 0 || kBM % HW == 0);
}
struct TmaConvHost {
  TmaConv tc;
  const void* x;
  int64_t frames;
  int H, W, C, convt;
};

static int g_num_sms = 0;
static int g_conv_wres = 1;     // npvp_set_option("conv_wres", 0): TMA-window convolutions stream their weights with every A tile again (A/B switch)
static int g_gemm_prefetch = 1; // npvp_set_option("gemm_prefetch", 0): no L2 prefetch of the A operand ahead of the TMA ring (A/B switch)
static int g_conv_tma = 1;      // npvp_set_option("conv_tma", 0): stride-1 zero-padded convolutions use the cp.async gather again (A/B switch)
static int g_use_2cta = -1;     // npvp_set_option("gemm_2cta", v): 1 = always for N >= 256, 0 = never, -1 (default) = when K >= 1024

// The direct epilogue (see epi_direct_chunk) needs whole 32-column chunks, 32-byte aligned output and
// residual rows, and alpha == 1 without a post-ReLU.
static int g_epi_direct = 1;   // npvp_set_option("gemm_epi_direct", 0): always transpose through shared memory (A/B switch, identical results)
static bool epi_direct_ok(const EpiParams& e, int64_t N) {
  if (!g_epi_direct || N % 32 != 0 || e.alpha != 1.0f || e.post_relu) return false;
  if (e.res1 && !e.res1_bf16) return false;      // fp32 residual rows: 128 B per thread in four sector requests measured slower (76.8 vs 56.3 us)
  if (e.out_bf16 && (e.ld_out % 16 != 0 || (uintptr_t)e.out_bf16 % 32 != 0)) return false;
  if (e.out_f32 && (e.ld_out % 8 != 0 || (uintptr_t)e.out_f32 % 32 != 0)) return false;
  if (e.res1 && (e.res1_bf16 ? (e.ld_res % 16 != 0) : (e.ld_res % 8 != 0))) return false;
  if (e.res2 && e.ld_res % 16 != 0) return false;
  if (((uintptr_t)e.res1 | (uintptr_t)e.res2) % 32 != 0) return false;
  return true;
}

template <int BN, int ACT, int RES, int OUT, int CONV>
static int launch_v2_inst(const CUtensorMap& ta, const CUtensorMap& tb, int64_t M, int64_t N, int64_t K, const EpiParams& e,
                          const ConvGather& cg, const TmaConv& tc, unsigned grid, cudaStream_t st) {
  using Cfg = Gemm2Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(gemm_tcgen05_v2_kernel<BN, ACT, RES, OUT, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (err != cudaSuccess) { npvp_set_error("cudaFuncSetAttribute(v2, smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(err)); return NPVP_ERR_CUDA; }
    attr_set = true;
  }
  gemm_tcgen05_v2_kernel<BN, ACT, RES, OUT, CONV><<<grid, CONV == 1 ? kGemm2Threads + 64 : kGemm2Threads, Cfg::kSmemBytes, st>>>(ta, tb, M, N, K, e, cg, tc);
  NPVP_LAUNCH_CHECK(CONV == 1 ? "gemm_tcgen05_v2_kernel<conv gather>" : (CONV ? "gemm_tcgen05_v2_kernel<conv tma>" : "gemm_tcgen05_v2_kernel"));
  return NPVP_OK;
}

// Epilogue specialisations used by the engines; anything else runs the run-time-flag instantiation <-1,-1,-1>.
template <int BN>
static int launch_tcgen05_v2(const void* A, int64_t lda, const void* W, int64_t ldw, int64_t M, int64_t N, int64_t K,
                             const EpiParams& e, cudaStream_t st, const ConvGather* conv = nullptr, const TmaConvHost* tmc = nullptr) {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  CUtensorMap ta, tb;
  int rc = make_tmap_2d(&tb, W, N, K, ldw, BN, e.fp16);
  if (rc) return rc;
  TmaConv tc0 = tmc ? tmc->tc : TmaConv{};
  tc0.pf = g_gemm_prefetch;
  if (tmc) {
    if ((rc = make_tmap_nhwc(&ta, tmc->x, tmc->frames, tmc->H, tmc->W, tmc->C, e.fp16))) return rc;
    // resident weights (see TmaConv): one n-tile, live W beside >= 3 A stages inside the stage ring's shared memory
    using Cfg = Gemm2Cfg<BN>;
    const int64_t live_rows = tmc->convt ? 9LL * tc0.cout : (int64_t)N * tc0.taps;          // summed over taps
    const int64_t wbytes = live_rows * tc0.kb_per_tap * 128;
    const int wbox = tmc->convt ? tc0.cout : (int)N;
    const int64_t ring = (int64_t)Cfg::kStages * Cfg::kStageBytes;
    tc0.wres = 0;
    if (g_conv_wres && N <= BN && N % 16 == 0 && wbox <= 256 && wbytes + 3 * Cfg::kABytes <= ring && wbytes < (1 << 20)) {
      tc0.wres = 1;
      tc0.wres_bytes = (int)wbytes;
      tc0.wbox = wbox;
      const int64_t nst = (ring - wbytes) / Cfg::kABytes;
      tc0.nst = (int)(nst < Cfg::kStages ? nst : Cfg::kStages);
      if ((rc = make_tmap_2d(&tb, W, N, K, ldw, wbox, e.fp16))) return rc;                   // boxes of `wbox` weight rows
    }
  }
  else if (conv) ta = tb;                              // unused in gather mode, keep the parameter valid
  else if ((rc = make_tmap_2d(&ta, A, M, K, lda, kBM, e.fp16))) return rc;
  const int64_t tiles = ceil_div64(M, kBM) * ceil_div64(N, BN);
  const unsigned grid = (unsigned)(tiles < g_num_sms ? tiles : g_num_sms);
  const ConvGather cg0 = conv ? *conv : ConvGather{};
  const int res = !e.res1 ? (e.res2 ? -1 : 0) : (!e.res2 ? (e.res1_bf16 ? 2 : 1) : ((e.res1_bf16 && e.res2_bf16) ? 3 : -1));
  const int out = (e.out_f32 && e.out_bf16) ? 2 : (e.out_f32 ? 1 : 0);
  const int act = e.act;
  if (tmc) {
    if (tmc->convt) {                                   // always the direct epilogue (the pixel-shuffled NHWC store exists only there)
      if (act == NPVP_ACT_RELU && res == 0 && out == 0) return launch_v2_inst<BN, NPVP_ACT_RELU, 0, kDirect, 3>(ta, tb, M, N, K, e, cg0, tc0, grid, st);
      if (act == NPVP_ACT_NONE && res == 0 && out == 0) return launch_v2_inst<BN, NPVP_ACT_NONE, 0, kDirect, 3>(ta, tb, M, N, K, e, cg0, tc0, grid, st);
      npvp_set_error("transposed conv: act NONE / RELU, 16-bit output without residuals only");
      return NPVP_ERR_INVALID;
    }
    const bool direct = epi_direct_ok(e, N);
#define NPVP_V2_TCONV(A_, R_, O_) if (act == A_ && res == R_ && out == O_) { \
      if (direct) return launch_v2_inst<BN, A_, R_, O_ + kDirect, 2>(ta, tb, M, N, K, e, cg0, tc0, grid, st); \
      return launch_v2_inst<BN, A_, R_, O_, 2>(ta, tb, M, N, K, e, cg0, tc0, grid, st); }
    NPVP_V2_TCONV(NPVP_ACT_RELU, 0, 0);
    NPVP_V2_TCONV(NPVP_ACT_RELU, 2, 0);
    NPVP_V2_TCONV(NPVP_ACT_NONE, 2, 0);
#undef NPVP_V2_TCONV
    return launch_v2_inst<BN, -1, -1, -1, 2>(ta, tb, M, N, K, e, cg0, tc0, grid, st);
  }
  const bool direct = epi_direct_ok(e, N);
  if (conv) {
#define NPVP_V2_CONV(A_, R_, O_) if (act == A_ && res == R_ && out == O_) { \
      if (direct) return launch_v2_inst<BN, A_, R_, O_ + kDirect, 1>(ta, tb, M, N, K, e, cg0, tc0, grid, st); \
      return launch_v2_inst<BN, A_, R_, O_, 1>(ta, tb, M, N, K, e, cg0, tc0, grid, st); }
    NPVP_V2_CONV(NPVP_ACT_RELU, 0, 0);   // conv / transposed conv + BN + ReLU
    NPVP_V2_CONV(NPVP_ACT_RELU, 2, 0);   // F3D conv: ReLU(BN(conv)) + x
    NPVP_V2_CONV(NPVP_ACT_NONE, 2, 0);   // ResnetBlock second conv + skip
    NPVP_V2_CONV(NPVP_ACT_NONE, 2, 1);   // ... last block: fp32 tokens out
#undef NPVP_V2_CONV
    return launch_v2_inst<BN, -1, -1, -1, 1>(ta, tb, M, N, K, e, cg0, tc0, grid, st);
  }
  if (e.frame_stats) {
    if (BN == 256 && act == NPVP_ACT_NONE && res == 0 && out == 0) {
      if (direct) return launch_v2_inst<256, NPVP_ACT_NONE, 0, kOutStats16 + kDirect, 0>(ta, tb, M, N, K, e, cg0, tc0, grid, st);
      return launch_v2_inst<256, NPVP_ACT_NONE, 0, kOutStats16, 0>(ta, tb, M, N, K, e, cg0, tc0, grid, st);
    }
    npvp_set_error("gemm: frame_stats needs the 256-wide tile, act NONE, no residual, 16-bit output only");
    return NPVP_ERR_INVALID;
  }
#define NPVP_V2_CASE(A_, R_, O_) if (act == A_ && res == R_ && out == O_) { \
    if (direct) return launch_v2_inst<BN, A_, R_, O_ + kDirect, 0>(ta, tb, M, N, K, e, cg0, tc0, grid, st); \
    return launch_v2_inst<BN, A_, R_, O_, 0>(ta, tb, M, N, K, e, cg0, tc0, grid, st); }
  NPVP_V2_CASE(NPVP_ACT_NONE, 0, 0);   // projections -> 16-bit
  NPVP_V2_CASE(NPVP_ACT_GELU, 0, 0);   // linear1
  NPVP_V2_CASE(NPVP_ACT_RELU, 0, 0);   // conv + BN + ReLU
  NPVP_V2_CASE(NPVP_ACT_NONE, 1, 1);   // out-proj / linear2: fp32 residual stream update
  NPVP_V2_CASE(NPVP_ACT_NONE, 0, 1);   // fc2 -> h3, latent heads
  NPVP_V2_CASE(NPVP_ACT_RELU, 2, 0);   // F3D conv: ReLU(BN(conv)) + x
  NPVP_V2_CASE(NPVP_ACT_RELU, 3, 0);   // non-local out_proj: y + gamma*ReLU(BN(.)) + x
  NPVP_V2_CASE(NPVP_ACT_NONE, 2, 0);   // ResnetBlock second conv + skip
  NPVP_V2_CASE(NPVP_ACT_NONE, 2, 1);   // ... last block: fp32 tokens out
#undef NPVP_V2_CASE
  return launch_v2_inst<BN, -1, -1, -1, 0>(ta, tb, M, N, K, e, cg0, tc0, grid, st);
}

template <int ACT, int RES, int OUT>
static int launch_2cta_inst(const CUtensorMap& ta, const CUtensorMap& tb, int64_t M, int64_t N, int64_t K, const EpiParams& e,
                            unsigned grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<ACT, RES, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm3Cfg::kSmemBytes);
    if (err != cudaSuccess) { npvp_set_error("cudaFuncSetAttribute(2cta, smem=%d): %s", Gemm3Cfg::kSmemBytes, cudaGetErrorString(err)); return NPVP_ERR_CUDA; }
    attr_set = true;
  }
  gemm_tcgen05_2cta_kernel<ACT, RES, OUT><<<grid, kGemm2Threads, Gemm3Cfg::kSmemBytes, st>>>(ta, tb, M, N, K, e);
  NPVP_LAUNCH_CHECK("gemm_tcgen05_2cta_kernel");
  return NPVP_OK;
}

static int launch_tcgen05_2cta(const void* A, int64_t lda, const void* W, int64_t ldw, int64_t M, int64_t N, int64_t K,
                               const EpiParams& e, cudaStream_t st) {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  CUtensorMap ta, tb;
  int rc = make_tmap_2d(&ta, A, M, K, lda, kBM, e.fp16);
  if (rc) return rc;
  rc = make_tmap_2d(&tb, W, N, K, ldw, Gemm3Cfg::BN / 2, e.fp16);
  if (rc) return rc;
  const int64_t tiles = ceil_div64(M, 2 * kBM) * ceil_div64(N, Gemm3Cfg::BN);
  const int64_t pairs = g_num_sms / 2;
  const unsigned grid = 2u * (unsigned)(tiles < pairs ? tiles : pairs);
  const int res = !e.res1 ? (e.res2 ? -1 : 0) : (!e.res2 ? (e.res1_bf16 ? 2 : 1) : ((e.res1_bf16 && e.res2_bf16) ? 3 : -1));
  const int out = (e.out_f32 && e.out_bf16) ? 2 : (e.out_f32 ? 1 : 0);
  const int act = e.act;
  const bool direct = epi_direct_ok(e, N);
#define NPVP_V3_CASE(A_, R_, O_) if (act == A_ && res == R_ && out == O_) { \
    if (direct) return launch_2cta_inst<A_, R_, O_ + kDirect>(ta, tb, M, N, K, e, grid, st); \
    return launch_2cta_inst<A_, R_, O_>(ta, tb, M, N, K, e, grid, st); }
  NPVP_V3_CASE(NPVP_ACT_NONE, 0, 0);
  NPVP_V3_CASE(NPVP_ACT_GELU, 0, 0);
  NPVP_V3_CASE(NPVP_ACT_RELU, 0, 0);
  NPVP_V3_CASE(NPVP_ACT_NONE, 1, 1);
  NPVP_V3_CASE(NPVP_ACT_NONE, 0, 1);
#undef NPVP_V3_CASE
  return launch_2cta_inst<-1, -1, -1>(ta, tb, M, N, K, e, grid, st);
}

static bool tma_compatible(const void* A, int64_t lda, const void* W, int64_t ldw, int64_t K) {
  return ((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && (lda % 8 == 0) && (ldw % 8 == 0) && (K % 8 == 0);
}

extern "C" int npvp_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int64_t M, int64_t N, int64_t K,
                              const npvp_epilogue_t* ep, int backend, void* stream) {
  NPVP_REQUIRE(A && W && ep, "npvp_gemm_bf16: null pointer");
  NPVP_REQUIRE(M > 0 && N > 0 && K > 0, "npvp_gemm_bf16: empty problem M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  NPVP_REQUIRE(lda >= K && ldw >= K, "npvp_gemm_bf16: leading dimension smaller than K");
  NPVP_REQUIRE(ep->out_f32 || ep->out_bf16, "npvp_gemm_bf16: no output buffer");
  NPVP_REQUIRE(ep->ld_out >= N, "npvp_gemm_bf16: ld_out < N");
  NPVP_REQUIRE(!(ep->res1 || ep->res2) || ep->ld_res >= N, "npvp_gemm_bf16: ld_res < N");
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int64_t kMaxRows = 65535LL * 64;   // grid.y limit of either backend
  if (M > kMaxRows) {                          // split along M: rows are independent
    for (int64_t m0 = 0; m0 < M; m0 += kMaxRows) {
      npvp_epilogue_t sub = *ep;
      const int64_t rows = (M - m0 < kMaxRows) ? (M - m0) : kMaxRows;
      if (sub.res1) sub.res1 = (const char*)sub.res1 + m0 * sub.ld_res * (sub.res1_bf16 ? 2 : 4);
      if (sub.res2) sub.res2 = (const char*)sub.res2 + m0 * sub.ld_res * (sub.res2_bf16 ? 2 : 4);
      if (sub.out_f32) sub.out_f32 = (char*)sub.out_f32 + m0 * sub.ld_out * 4;
      if (sub.out_bf16) sub.out_bf16 = (char*)sub.out_bf16 + m0 * sub.ld_out * 2;
      if (sub.frame_stats) sub.frame_stats += (m0 / 64) * (N / 256) * 4 * 2;
      int rc = npvp_gemm_bf16((const char*)A + m0 * lda * 2, lda, W, ldw, rows, N, K, &sub, backend, stream);
      if (rc) return rc;
    }
    return NPVP_OK;
  }
  EpiParams e = make_epi(ep);
  if (ep->frame_stats) {                        // fused per-frame statistics: one fixed route (persistent 1-CTA kernel, 256-wide tiles)
    NPVP_REQUIRE(M % 64 == 0 && N % 256 == 0 && ep->out_bf16 && !ep->out_f32 && !ep->res1 && !ep->res2 && ep->act == NPVP_ACT_NONE,
                 "npvp_gemm_bf16: frame_stats needs M %% 64 == 0, N %% 256 == 0, 16-bit output only, act NONE, no residuals");
    NPVP_REQUIRE(backend == NPVP_GEMM_AUTO || backend == NPVP_GEMM_TCGEN05, "npvp_gemm_bf16: frame_stats is implemented by the default back-end only");
    NPVP_REQUIRE(tma_compatible(A, lda, W, ldw, K) && ep->ld_out % 8 == 0 && (uintptr_t)ep->out_bf16 % 16 == 0,
                 "npvp_gemm_bf16: frame_stats needs TMA-compatible operands");
    return launch_tcgen05_v2<256>(A, lda, W, ldw, M, N, K, e, st);
  }
  const bool vec_ok = (ep->ld_out % 8 == 0) && (!ep->out_f32 || (uintptr_t)ep->out_f32 % 16 == 0) &&
                      (!ep->out_bf16 || (uintptr_t)ep->out_bf16 % 16 == 0);
  // AUTO: tensor path whenever the operands are TMA-expressible.  The choice must not depend on M (the batch), otherwise a
  // clip would be computed differently alone and inside a batch; TMA zero-fills boxes that overhang small operands.
  if (backend == NPVP_GEMM_AUTO)
    backend = (tma_compatible(A, lda, W, ldw, K) && vec_ok && N % 4 == 0 &&
               (!(ep->res1 || ep->res2) || (ep->ld_res % 4 == 0 && (uintptr_t)ep->res1 % 16 == 0 && (uintptr_t)ep->res2 % 16 == 0)))
                  ? NPVP_GEMM_TCGEN05 : NPVP_GEMM_SIMT;   // K < 64: TMA zero-fills the k-block
  if (backend == NPVP_GEMM_TCGEN05 || backend == NPVP_GEMM_TCGEN05_2CTA) {
    NPVP_REQUIRE(tma_compatible(A, lda, W, ldw, K) && vec_ok, "npvp_gemm_bf16: operands not 16-byte aligned / K,ld not multiples of 8 for the TMA path");
    const bool res_ok = (!(ep->res1 || ep->res2)) || (ep->ld_res % 4 == 0 && (uintptr_t)ep->res1 % 16 == 0 && (uintptr_t)ep->res2 % 16 == 0);
    if (backend == NPVP_GEMM_TCGEN05_2CTA) {
      NPVP_REQUIRE(N % 4 == 0 && res_ok, "npvp_gemm_bf16: 2-CTA path needs N %% 4 == 0 and 16-byte aligned residuals");
      return launch_tcgen05_2cta(A, lda, W, ldw, M, N, K, e, st);
    }
    NPVP_REQUIRE(N % 4 == 0 && res_ok, "npvp_gemm_bf16: the tcgen05 kernels need N %% 4 == 0 and 16-byte aligned residuals (use NPVP_GEMM_SIMT)");
    {                                                                 // persistent kernel; tile width never depends on M
      // 2-CTA (cta_group::2) tiles pay off once the main loop dominates the tile: measured 1150 vs 1011 TFLOP/s at K = 2048,
      // 912 vs 1013 at K = 512 (M = 40960).  The choice depends on N and K only, never on M (batch invariance).
      if (N >= 256 && (g_use_2cta == 1 || (g_use_2cta < 0 && K >= 1024))) return launch_tcgen05_2cta(A, lda, W, ldw, M, N, K, e, st);
      if (N >= 256) return launch_tcgen05_v2<256>(A, lda, W, ldw, M, N, K, e, st);
      if (N > 64) return launch_tcgen05_v2<128>(A, lda, W, ldw, M, N, K, e, st);
      return launch_tcgen05_v2<64>(A, lda, W, ldw, M, N, K, e, st);
    }
  }
  NPVP_REQUIRE(backend == NPVP_GEMM_SIMT, "npvp_gemm_bf16: unknown backend %d", backend);
  dim3 grid((unsigned)ceil_div64(N, 64), (unsigned)ceil_div64(M, 64));
  gemm_simt_kernel<h16><<<grid, 256, 0, st>>>((const h16*)A, lda, (const h16*)W, ldw, M, N, K, e);
  NPVP_LAUNCH_CHECK("gemm_simt_kernel<h16>");
  return NPVP_OK;
}

extern "C" int npvp_gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, int64_t M, int64_t N, int64_t K,
                             const float* bias, int act, float* out, int64_t ldo, void* stream) {
  NPVP_REQUIRE(A && W && out, "npvp_gemm_f32: null pointer");
  NPVP_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldo >= N, "npvp_gemm_f32: bad shape");
  EpiParams e = {};
  e.bias = bias;
  e.out_f32 = out;
  e.alpha = 1.0f;
  e.act = act;
  e.ld_out = ldo;
  dim3 grid((unsigned)ceil_div64(N, 64), (unsigned)ceil_div64(M, 64));
  gemm_simt_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, W, ldw, M, N, K, e);
  NPVP_LAUNCH_CHECK("gemm_simt_kernel<float>");
  return NPVP_OK;
}


// Convolution as an implicit GEMM: out[(f,oy,ox), n] = epilogue( sum_{ky,kx,ci} x[f, oy*stride-pad+ky, ox*stride-pad+kx, ci] * W[n, (ky,kx,ci)] ).
extern "C" int npvp_conv_gemm_bf16(const void* x, int64_t frames, int H, int W, int C, int KH, int KW, int stride, int pad,
                                   int pad_mode, int Ho, int Wo, int phase_major, const void* Wt, int64_t ldw, int64_t N,
                                   const npvp_epilogue_t* ep, void* stream) {
  NPVP_REQUIRE(x && Wt && ep && frames > 0, "npvp_conv_gemm_bf16: null pointer");
  NPVP_REQUIRE(C % 32 == 0 && (C < 64 || C % 64 == 0), "npvp_conv_gemm_bf16: C must be 32 or a multiple of 64 (got %d)", C);
  NPVP_REQUIRE(KH > 0 && KW > 0 && stride > 0 && Ho > 0 && Wo > 0 && N > 0, "npvp_conv_gemm_bf16: bad geometry");
  NPVP_REQUIRE(!phase_major || (H % 2 == 0 && W % 2 == 0), "npvp_conv_gemm_bf16: phase-major input needs even H, W");
  NPVP_REQUIRE(pad_mode == NPVP_PAD_ZERO || (pad < H && pad < W), "npvp_conv_gemm_bf16: reflect/replicate pad must be smaller than the image");
  const int64_t M = frames * Ho * Wo, K = (int64_t)KH * KW * C;
  NPVP_REQUIRE(ldw >= K && ldw % 8 == 0 && (uintptr_t)Wt % 16 == 0 && (uintptr_t)x % 16 == 0, "npvp_conv_gemm_bf16: weight/activation alignment");
  NPVP_REQUIRE(ep->out_f32 || ep->out_bf16, "npvp_conv_gemm_bf16: no output buffer");
  NPVP_REQUIRE(ep->ld_out >= N && ep->ld_out % 8 == 0 && N % 4 == 0, "npvp_conv_gemm_bf16: ld_out >= N, ld_out %% 8 == 0, N %% 4 == 0 required");
  NPVP_REQUIRE((!ep->out_f32 || (uintptr_t)ep->out_f32 % 16 == 0) && (!ep->out_bf16 || (uintptr_t)ep->out_bf16 % 16 == 0), "npvp_conv_gemm_bf16: output alignment");
  NPVP_REQUIRE(!(ep->res1 || ep->res2) || (ep->ld_res >= N && ep->ld_res % 4 == 0 && (uintptr_t)ep->res1 % 16 == 0 && (uintptr_t)ep->res2 % 16 == 0),
               "npvp_conv_gemm_bf16: residual alignment");
  ConvGather cg;
  cg.x = (const h16*)x;
  cg.H = H; cg.W = W; cg.C = C; cg.KH = KH; cg.KW = KW; cg.stride = stride; cg.pad = pad; cg.pad_mode = pad_mode;
  cg.Ho = Ho; cg.Wo = Wo; cg.phase_major = phase_major;
  EpiParams e = make_epi(ep);
  cudaStream_t st = (cudaStream_t)stream;
  NPVP_REQUIRE((int64_t)frames * H * W < 0xffffffffll && M < 0xffffffffll, "npvp_conv_gemm_bf16: more than 2^32 - 1 input or output pixels per launch");
  if (g_conv_tma && stride == 1 && !phase_major && Ho == H && Wo == W && KH == 2 * pad + 1 && KW == 2 * pad + 1 &&
      (pad_mode == NPVP_PAD_ZERO || pad == 0) && tma_conv_geometry(H, W, C)) {
    // stride-1 "same" convolution with zero padding: shifted-window A tiles by TMA (no gather warps)
    TmaConvHost h;
    h.tc.taps = KH * KW; h.tc.KW = KW; h.tc.pad = pad; h.tc.kb_per_tap = C / kBK; h.tc.HW = H * W; h.tc.W = W; h.tc.cout = 0;
    h.tc.wres = 0; h.tc.wres_bytes = 0; h.tc.nst = 0; h.tc.wbox = 0;
    h.x = x; h.frames = frames; h.H = H; h.W = W; h.C = C; h.convt = 0;
    if (N >= 256) return launch_tcgen05_v2<256>(nullptr, 0, Wt, ldw, M, N, K, e, st, nullptr, &h);
    if (N > 64) return launch_tcgen05_v2<128>(nullptr, 0, Wt, ldw, M, N, K, e, st, nullptr, &h);
    return launch_tcgen05_v2<64>(nullptr, 0, Wt, ldw, M, N, K, e, st, nullptr, &h);
  }
  if (N >= 256 && C >= 64) return launch_tcgen05_v2<256>(nullptr, 0, Wt, ldw, M, N, K, e, st, &cg);   // (C == 32 needs the 2-tap table)
  if (N > 64) return launch_tcgen05_v2<128>(nullptr, 0, Wt, ldw, M, N, K, e, st, &cg);
  return launch_tcgen05_v2<64>(nullptr, 0, Wt, ldw, M, N, K, e, st, &cg);
}


// Transposed convolution ConvTranspose2d(3, stride 2, pad 1, output_pad 1) as ONE implicit GEMM over the 2x2 input
// neighbourhood: out[f, 2y+py, 2x+px, co] = sum_{dy,dx,ci} x[f, y+dy, x+dx, ci] * Wt[(q(py,px), co), (dy,dx,ci)], phases ordered
// q = (0,0), (0,1), (1,1), (1,0).  Only the 9 live (phase, tap) blocks are loaded and multiplied; x and out are plain NHWC.
extern "C" int npvp_convt_gemm_bf16(const void* x, int64_t frames, int H, int W, int Cin, const void* Wt, int64_t ldw, int Cout,
                                    const npvp_epilogue_t* ep, void* stream) {
  NPVP_REQUIRE(x && Wt && ep && frames > 0, "npvp_convt_gemm_bf16: null pointer");
  NPVP_REQUIRE(H > 0 && W > 0 && (W & (W - 1)) == 0 && tma_conv_geometry(H, W, Cin),
               "npvp_convt_gemm_bf16: needs Cin %% 64 == 0, W a power of two <= 128 and frames tiled by 128-pixel blocks (H=%d W=%d Cin=%d)", H, W, Cin);
  NPVP_REQUIRE(Cout > 0 && Cout % 32 == 0, "npvp_convt_gemm_bf16: Cout must be a multiple of 32 (got %d)", Cout);
  const int64_t M = frames * H * W, N = 4LL * Cout, K = 4LL * Cin;
  NPVP_REQUIRE(ldw >= K && ldw % 8 == 0 && (uintptr_t)Wt % 16 == 0 && (uintptr_t)x % 16 == 0, "npvp_convt_gemm_bf16: weight/activation alignment");
  NPVP_REQUIRE(ep->out_bf16 && !ep->out_f32 && !ep->res1 && !ep->res2 && !ep->frame_stats && (uintptr_t)ep->out_bf16 % 16 == 0,
               "npvp_convt_gemm_bf16: one 16-byte aligned 16-bit output, no residuals");
  NPVP_REQUIRE(4 * M < 0xffffffffll, "npvp_convt_gemm_bf16: more than 2^32 - 1 output pixels per launch");
  NPVP_REQUIRE(ep->alpha == 1.0f && !ep->post_relu, "npvp_convt_gemm_bf16: alpha must be 1 and post_relu 0");
  NPVP_REQUIRE((uintptr_t)ep->out_bf16 % 32 == 0, "npvp_convt_gemm_bf16: output must be 32-byte aligned (sector stores)");
  EpiParams e = make_epi(ep);
  TmaConvHost h;
  h.tc.taps = 4; h.tc.KW = 2; h.tc.pad = 0; h.tc.kb_per_tap = Cin / kBK; h.tc.HW = H * W; h.tc.W = W; h.tc.cout = Cout;
  h.tc.wres = 0; h.tc.wres_bytes = 0; h.tc.nst = 0; h.tc.wbox = 0;
  h.x = x; h.frames = frames; h.H = H; h.W = W; h.C = Cin; h.convt = 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (N >= 256) return launch_tcgen05_v2<256>(nullptr, 0, Wt, ldw, M, N, K, e, st, nullptr, &h);
  return launch_tcgen05_v2<128>(nullptr, 0, Wt, ldw, M, N, K, e, st, nullptr, &h);
}

void npvp_set_ffn_scalar(int v);   // predictor_kernels.cu
void npvp_set_ffn_mid16_mode(int m);   // ffn_mid16.cu
void npvp_head_tc_enable(int v);       // head_tc.cu
void npvp_head_tc_wait(int v);         // head_tc.cu
void npvp_stem_tc_enable(int v);       // head_tc.cu

extern "C" int npvp_set_option(const char* name, int value) {
  NPVP_REQUIRE(name != nullptr, "npvp_set_option: null name");
  if (strcmp(name, "gemm_2cta") == 0) { g_use_2cta = value; return NPVP_OK; }
  if (strcmp(name, "gemm_epi_direct") == 0) { g_epi_direct = value; return NPVP_OK; }
  if (strcmp(name, "conv_tma") == 0) { g_conv_tma = value; return NPVP_OK; }
  if (strcmp(name, "gemm_prefetch") == 0) { g_gemm_prefetch = value; return NPVP_OK; }
  if (strcmp(name, "conv_wres") == 0) { g_conv_wres = value; return NPVP_OK; }
  if (strcmp(name, "ffn_scalar") == 0) { npvp_set_ffn_scalar(value); return NPVP_OK; }
  if (strcmp(name, "ffn_mid16_mode") == 0) { npvp_set_ffn_mid16_mode(value); return NPVP_OK; }
  if (strcmp(name, "head_tc") == 0) { npvp_head_tc_enable(value); return NPVP_OK; }
  if (strcmp(name, "head_tc_wait") == 0) { npvp_head_tc_wait(value); return NPVP_OK; }
  if (strcmp(name, "stem_tc") == 0) { npvp_stem_tc_enable(value); return NPVP_OK; }
  NPVP_REQUIRE(false, "npvp_set_option: unknown option '%s'", name);
  return NPVP_ERR_INVALID;
}
