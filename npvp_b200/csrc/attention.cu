// Attention cores of the factorised VidHRFormer attention (8 heads x 64): 4x4 spatial windows (L = 16) and
// per-pixel temporal / encoder-decoder sequences (L = T <= 32).  The window / time gathers are folded into the row
// index math, so no permuted copies are materialised.  HBM-bound by design: Q, K, V are read once, O written once.
//
// One warp owns one (sequence, head).  The sequences are far too short for tcgen05 tiles (a 128-row UMMA would be
// >87 % padding), so the two tiny contractions use warp-level mma.sync.m16n8k16 and the softmax stays in registers:
//   S = (Q K^T) / 8      bf16 x bf16 -> fp32, A/B fragments loaded straight from global (K rows are B-fragment friendly)
//   P = softmax(S)       fp32 in the accumulator layout, row max / sum via quad shuffles, masks as predicates
//   O = P V              P re-used as an fp16 A fragment (p in [0,1]: 11-bit mantissa), V staged to smem as exact fp16
//                        copies of the bf16 values and fetched with ldmatrix.trans
// r01 profile: the first SIMT version was instruction-bound (~2.5k warp-instructions per (sequence, head)); this
// formulation needs ~200.
#include "common.cuh"

constexpr int kHeads = 8;
constexpr int kHd = 64;
constexpr int kMaxL = 32;
constexpr int kWarpsPerBlock = 4;
constexpr int kVStride = kHd + 8;   // halves per staged V row: 144 B keeps the 8 ldmatrix row addresses on distinct banks

// Read-only 16-byte load.
__device__ __forceinline__ uint4 ldg128_batched(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

template <int MODE>
__device__ __forceinline__ int64_t seq_row(int64_t s, int i, int T) {
  if (MODE == NPVP_ATTN_SPATIAL_WINDOW) {
    const int64_t f = s >> 2;
    const int qh = (int)((s >> 1) & 1), qw = (int)(s & 1);
    return f * 64 + (4 * qh + (i >> 2)) * 8 + 4 * qw + (i & 3);
  } else {
    const int64_t n = s >> 6;
    const int hw = (int)(s & 63);
    return (n * T + i) * 64 + hw;
  }
}

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// two bf16 (packed) -> two fp16 (packed); exact for |x| in [6.1e-5, 65504] (bf16 has fewer mantissa bits than fp16)
__device__ __forceinline__ uint32_t bf16x2_to_f16x2(uint32_t u) {
  const float2 f = unpack_bf16x2(u);
  return pack_f16x2(f.x, f.y);
}

// KT = key tiles of 8 held in registers: 2 for Lk <= 16 (every shipped configuration: windows of 16, T <= 12), 4 up to kMaxL.
template <int MODE, int KT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
attention_kernel(const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k, int64_t ldk, const bf16* __restrict__ v,
                 int64_t ldv, bf16* __restrict__ out, int64_t ldo, int Tq, int Tk, int mask_last) {
  __shared__ __align__(16) __half sV[kWarpsPerBlock][KT * 8][kVStride];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int64_t s = blockIdx.x;
  const int head = blockIdx.y * kWarpsPerBlock + w;
  const int Lq = (MODE == NPVP_ATTN_SPATIAL_WINDOW) ? 16 : Tq;
  const int Lk = (MODE == NPVP_ATTN_SPATIAL_WINDOW) ? 16 : Tk;
  const int col = head * kHd;
  const int n_kt = (Lk + 7) >> 3;        // key tiles of 8 (S columns)
  const int n_kk = (Lk + 15) >> 4;       // key tiles of 16 (P V reduction steps)

  // A warp lives for one (sequence, head): ~5 KB of traffic, so its run time is the number of dependent global round trips
  // and of memory instructions (the first version issued 48 4-byte loads and 16 4-byte stores per thread, half of every
  // store sector unused, and ptxas sank the loads next to the mma that consumes them: 3 TB/s).  Every access is now 16 bytes:
  //  * the contraction over the 64 head dims is a sum, so Q and K may enumerate the dims in any common order: thread `tig`
  //    takes the 32 contiguous bytes [16 tig, 16 tig + 16) of a row and feeds dims 16 tig + 4 kk + {0,1} / + {2,3} to the
  //    k-slots (2 tig, 2 tig + 1) / (2 tig + 8, 2 tig + 9) of step kk: 2 x LDG.128 per row instead of 8 x LDG.32;
  //  * V is staged with its columns permuted (physical dim 16 t + 2 d + e -> column 8 d + 2 t + e), so the PV accumulators of
  //    thread `tig` are the physical dims [16 tig, 16 tig + 16) of its two rows: 2 x STG.128 per row, full sectors.
  // ---- V rows (keys): lane = (row 4 i + (lane >> 3), 16-byte chunk lane & 7) ----
  const int vc = lane & 7, vr = lane >> 3;
  uint4 vraw[KT * 2];
#pragma unroll
  for (int i = 0; i < KT * 2; ++i) {
    const int j = 4 * i + vr;
    vraw[i] = (j < Lk) ? ldg128_batched(reinterpret_cast<const uint4*>(v + seq_row<MODE>(s, j, Tk) * ldv + col) + vc) : make_uint4(0u, 0u, 0u, 0u);
  }

  // ---- K as B fragments: key tile t (8 keys = n), step kk: b0 = kr[2 kk], b1 = kr[2 kk + 1] of the thread's 32 bytes ----
  uint32_t kb[KT][4][2];
#pragma unroll
  for (int t = 0; t < KT; ++t) {
    const int j = 8 * t + gid;
    const bool ok = (t < n_kt) && (j < Lk);
    const uint4* krow = reinterpret_cast<const uint4*>(k + (ok ? seq_row<MODE>(s, j, Tk) : 0) * ldk + col) + 2 * tig;
    const uint4 lo = ok ? ldg128_batched(krow) : make_uint4(0u, 0u, 0u, 0u), hi = ok ? ldg128_batched(krow + 1) : make_uint4(0u, 0u, 0u, 0u);
    kb[t][0][0] = lo.x; kb[t][0][1] = lo.y; kb[t][1][0] = lo.z; kb[t][1][1] = lo.w;
    kb[t][2][0] = hi.x; kb[t][2][1] = hi.y; kb[t][3][0] = hi.z; kb[t][3][1] = hi.w;
  }

  // ---- query block: a0 / a2 of step kk = the same two registers of row gid, a1 / a3 of row gid + 8 ----
  uint32_t qa[4][4];
  auto load_q = [&](int m0) {
    const int r0 = m0 + gid, r1 = m0 + gid + 8;
    const bool ok0 = r0 < Lq, ok1 = r1 < Lq;
    const uint4* q0 = reinterpret_cast<const uint4*>(q + seq_row<MODE>(s, ok0 ? r0 : 0, Tq) * ldq + col) + 2 * tig;
    const uint4* q1 = reinterpret_cast<const uint4*>(q + seq_row<MODE>(s, ok1 ? r1 : 0, Tq) * ldq + col) + 2 * tig;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const uint4 a_lo = ok0 ? ldg128_batched(q0) : z, a_hi = ok0 ? ldg128_batched(q0 + 1) : z;
    const uint4 b_lo = ok1 ? ldg128_batched(q1) : z, b_hi = ok1 ? ldg128_batched(q1 + 1) : z;
    qa[0][0] = a_lo.x; qa[0][2] = a_lo.y; qa[1][0] = a_lo.z; qa[1][2] = a_lo.w;
    qa[2][0] = a_hi.x; qa[2][2] = a_hi.y; qa[3][0] = a_hi.z; qa[3][2] = a_hi.w;
    qa[0][1] = b_lo.x; qa[0][3] = b_lo.y; qa[1][1] = b_lo.z; qa[1][3] = b_lo.w;
    qa[2][1] = b_hi.x; qa[2][3] = b_hi.y; qa[3][1] = b_hi.z; qa[3][3] = b_hi.w;
  };
  load_q(0);

  // ---- stage V as fp16 (exact copies of the bf16 values) with permuted columns; rows >= Lk are zero ----
  {
    const int cbase = 8 * ((vc & 1) * 4) + 2 * (vc >> 1);      // register i2 of the chunk -> columns cbase + 8 i2, + 1
#pragma unroll
    for (int i = 0; i < KT * 2; ++i) {
      __half* row = &sV[w][4 * i + vr][cbase];
      *reinterpret_cast<uint32_t*>(row) = bf16x2_to_f16x2(vraw[i].x);
      *reinterpret_cast<uint32_t*>(row + 8) = bf16x2_to_f16x2(vraw[i].y);
      *reinterpret_cast<uint32_t*>(row + 16) = bf16x2_to_f16x2(vraw[i].z);
      *reinterpret_cast<uint32_t*>(row + 24) = bf16x2_to_f16x2(vraw[i].w);
    }
  }
  __syncwarp();

  for (int m0 = 0; m0 < Lq; m0 += 16) {
    const int r0 = m0 + gid, r1 = m0 + gid + 8;             // the two query rows this thread holds
    const bool ok0 = r0 < Lq, ok1 = r1 < Lq;
    const int64_t row0 = seq_row<MODE>(s, ok0 ? r0 : 0, Tq), row1 = seq_row<MODE>(s, ok1 ? r1 : 0, Tq);
    // ---- S = Q K^T ----
    float sc[KT][4];
#pragma unroll
    for (int t = 0; t < KT; ++t) { sc[t][0] = sc[t][1] = sc[t][2] = sc[t][3] = 0.f; }
    if (m0 > 0) load_q(m0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int t = 0; t < KT; ++t)
        if (t < n_kt) mma_bf16_16816(sc[t], qa[kk], kb[t][kk][0], kb[t][kk][1]);
    }
    // ---- softmax over keys (thread holds cols 8t+2tig, +1 of rows r0 (c0,c1) and r1 (c2,c3)) ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int t = 0; t < KT; ++t)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 8 * t + 2 * tig + e;
        const bool dead = (t >= n_kt) || (j >= Lk);
        const bool m0k = dead || (mask_last && j == Lk - 1 && r0 < Lq - 1);
        const bool m1k = dead || (mask_last && j == Lk - 1 && r1 < Lq - 1);
        sc[t][e] = m0k ? -INFINITY : sc[t][e] * 0.125f;
        sc[t][2 + e] = m1k ? -INFINITY : sc[t][2 + e] * 0.125f;
        mx0 = fmaxf(mx0, sc[t][e]);
        mx1 = fmaxf(mx1, sc[t][2 + e]);
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int t = 0; t < KT; ++t)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float p0 = (sc[t][e] == -INFINITY) ? 0.f : expf(sc[t][e] - mx0);
        const float p1 = (sc[t][2 + e] == -INFINITY) ? 0.f : expf(sc[t][2 + e] - mx1);
        sc[t][e] = p0;
        sc[t][2 + e] = p1;
        sum0 += p0;
        sum1 += p1;
      }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;     // every live row has at least one unmasked key
    // ---- O = P V: P (fp16) as A fragments from the accumulator layout; V^T fragments via ldmatrix.trans ----
    float o[8][4];
#pragma unroll
    for (int d = 0; d < 8; ++d) { o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < KT / 2; ++kk) {
      if (kk < n_kk) {
        uint32_t a[4];
        a[0] = pack_f16x2(sc[2 * kk][0] * inv0, sc[2 * kk][1] * inv0);
        a[1] = pack_f16x2(sc[2 * kk][2] * inv1, sc[2 * kk][3] * inv1);
        a[2] = pack_f16x2(sc[2 * kk + 1][0] * inv0, sc[2 * kk + 1][1] * inv0);
        a[3] = pack_f16x2(sc[2 * kk + 1][2] * inv1, sc[2 * kk + 1][3] * inv1);
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          uint32_t b0, b1;     // lanes 0-15 supply the row addresses: keys 16kk + (lane & 15), dims 8d..8d+7
          ldmatrix_x2_trans(b0, b1, &sV[w][16 * kk + (lane & 15)][8 * d]);
          mma_f16_16816(o[d], a, b0, b1);
        }
      }
    }
    // ---- store: the thread's 16 accumulator columns of a row are the physical dims 16 tig + 2 d + {0,1}: 32 contiguous bytes ----
    if (ok0) {
      uint4* dst = reinterpret_cast<uint4*>(out + row0 * ldo + col) + 2 * tig;
      dst[0] = make_uint4(pack_bf16x2(o[0][0], o[0][1]), pack_bf16x2(o[1][0], o[1][1]), pack_bf16x2(o[2][0], o[2][1]), pack_bf16x2(o[3][0], o[3][1]));
      dst[1] = make_uint4(pack_bf16x2(o[4][0], o[4][1]), pack_bf16x2(o[5][0], o[5][1]), pack_bf16x2(o[6][0], o[6][1]), pack_bf16x2(o[7][0], o[7][1]));
    }
    if (ok1) {
      uint4* dst = reinterpret_cast<uint4*>(out + row1 * ldo + col) + 2 * tig;
      dst[0] = make_uint4(pack_bf16x2(o[0][2], o[0][3]), pack_bf16x2(o[1][2], o[1][3]), pack_bf16x2(o[2][2], o[2][3]), pack_bf16x2(o[3][2], o[3][3]));
      dst[1] = make_uint4(pack_bf16x2(o[4][2], o[4][3]), pack_bf16x2(o[5][2], o[5][3]), pack_bf16x2(o[6][2], o[6][3]), pack_bf16x2(o[7][2], o[7][3]));
    }
  }
}

extern "C" int npvp_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                              int64_t ldo, int mode, int64_t n_clips, int Tq, int Tk, int mask_last, void* stream) {
  NPVP_REQUIRE(q && k && v && out, "npvp_attention: null pointer");
  NPVP_REQUIRE(n_clips > 0 && Tq > 0 && Tk > 0, "npvp_attention: empty input");
  NPVP_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "npvp_attention: row strides must be multiples of 8");
  NPVP_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) % 16 == 0, "npvp_attention: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 block(kWarpsPerBlock * 32);
  if (mode == NPVP_ATTN_SPATIAL_WINDOW) {
    NPVP_REQUIRE(Tq == Tk, "npvp_attention: spatial window attention needs Tq == Tk");
    const int64_t nseq = n_clips * Tq * 4;
    attention_kernel<NPVP_ATTN_SPATIAL_WINDOW, 2><<<dim3((unsigned)nseq, kHeads / kWarpsPerBlock), block, 0, st>>>(
        (const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (bf16*)out, ldo, Tq, Tk, 0);
  } else if (mode == NPVP_ATTN_TEMPORAL) {
    NPVP_REQUIRE(Tq <= kMaxL && Tk <= kMaxL, "npvp_attention: temporal length above %d not supported (Tq=%d Tk=%d)", kMaxL, Tq, Tk);
    const int64_t nseq = n_clips * 64;
    if (Tk <= 16)
      attention_kernel<NPVP_ATTN_TEMPORAL, 2><<<dim3((unsigned)nseq, kHeads / kWarpsPerBlock), block, 0, st>>>(
          (const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (bf16*)out, ldo, Tq, Tk, mask_last);
    else
      attention_kernel<NPVP_ATTN_TEMPORAL, 4><<<dim3((unsigned)nseq, kHeads / kWarpsPerBlock), block, 0, st>>>(
          (const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (bf16*)out, ldo, Tq, Tk, mask_last);
  } else {
    NPVP_REQUIRE(false, "npvp_attention: unknown mode %d", mode);
  }
  NPVP_LAUNCH_CHECK("attention_kernel");
  return NPVP_OK;
}
