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

template <int MODE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
attention_kernel(const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k, int64_t ldk, const bf16* __restrict__ v,
                 int64_t ldv, bf16* __restrict__ out, int64_t ldo, int Tq, int Tk, int mask_last) {
  __shared__ __align__(16) __half sV[kWarpsPerBlock][kMaxL][kVStride];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int64_t s = blockIdx.x;
  const int head = blockIdx.y * kWarpsPerBlock + w;
  const int Lq = (MODE == NPVP_ATTN_SPATIAL_WINDOW) ? 16 : Tq;
  const int Lk = (MODE == NPVP_ATTN_SPATIAL_WINDOW) ? 16 : Tk;
  const int col = head * kHd;
  const int n_kt = (Lk + 7) >> 3;        // key tiles of 8 (S columns)
  const int n_kk = (Lk + 15) >> 4;       // key tiles of 16 (P V reduction steps)

  // ---- stage V (rows = keys) as fp16, zero-padded to a multiple of 16 keys ----
  for (int j = 0; j < n_kk * 16; ++j) {
    uint32_t u = 0u;
    if (j < Lk) u = bf16x2_to_f16x2(__ldg(reinterpret_cast<const uint32_t*>(v + seq_row<MODE>(s, j, Tk) * ldv + col) + lane));
    reinterpret_cast<uint32_t*>(&sV[w][j][0])[lane] = u;
  }
  __syncwarp();

  // ---- K as B fragments: key tile t (8 keys), d-step kk (16 dims): b0 = K[8t+gid][16kk+2tig..], b1 = ... +8 ----
  uint32_t kb[4][4][2];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int j = 8 * t + gid;
    const bool ok = (t < n_kt) && (j < Lk);
    const uint32_t* krow = reinterpret_cast<const uint32_t*>(k + (ok ? seq_row<MODE>(s, j, Tk) : 0) * ldk + col);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      kb[t][kk][0] = ok ? __ldg(krow + 8 * kk + tig) : 0u;
      kb[t][kk][1] = ok ? __ldg(krow + 8 * kk + 4 + tig) : 0u;
    }
  }

  for (int m0 = 0; m0 < Lq; m0 += 16) {
    const int r0 = m0 + gid, r1 = m0 + gid + 8;             // the two query rows this thread holds
    const bool ok0 = r0 < Lq, ok1 = r1 < Lq;
    const int64_t row0 = seq_row<MODE>(s, ok0 ? r0 : 0, Tq), row1 = seq_row<MODE>(s, ok1 ? r1 : 0, Tq);
    const uint32_t* q0 = reinterpret_cast<const uint32_t*>(q + row0 * ldq + col);
    const uint32_t* q1 = reinterpret_cast<const uint32_t*>(q + row1 * ldq + col);
    // ---- S = Q K^T ----
    float sc[4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t) { sc[t][0] = sc[t][1] = sc[t][2] = sc[t][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      a[0] = ok0 ? __ldg(q0 + 8 * kk + tig) : 0u;
      a[1] = ok1 ? __ldg(q1 + 8 * kk + tig) : 0u;
      a[2] = ok0 ? __ldg(q0 + 8 * kk + 4 + tig) : 0u;
      a[3] = ok1 ? __ldg(q1 + 8 * kk + 4 + tig) : 0u;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (t < n_kt) mma_bf16_16816(sc[t], a, kb[t][kk][0], kb[t][kk][1]);
    }
    // ---- softmax over keys (thread holds cols 8t+2tig, +1 of rows r0 (c0,c1) and r1 (c2,c3)) ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int t = 0; t < 4; ++t)
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
    for (int t = 0; t < 4; ++t)
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
    for (int kk = 0; kk < 2; ++kk) {
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
    // ---- store: thread holds dims 8d+2tig, +1 of rows r0 / r1 ----
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      if (ok0) reinterpret_cast<uint32_t*>(out + row0 * ldo + col)[4 * d + tig] = pack_bf16x2(o[d][0], o[d][1]);
      if (ok1) reinterpret_cast<uint32_t*>(out + row1 * ldo + col)[4 * d + tig] = pack_bf16x2(o[d][2], o[d][3]);
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
    attention_kernel<NPVP_ATTN_SPATIAL_WINDOW><<<dim3((unsigned)nseq, kHeads / kWarpsPerBlock), block, 0, st>>>(
        (const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (bf16*)out, ldo, Tq, Tk, 0);
  } else if (mode == NPVP_ATTN_TEMPORAL) {
    NPVP_REQUIRE(Tq <= kMaxL && Tk <= kMaxL, "npvp_attention: temporal length above %d not supported (Tq=%d Tk=%d)", kMaxL, Tq, Tk);
    const int64_t nseq = n_clips * 64;
    attention_kernel<NPVP_ATTN_TEMPORAL><<<dim3((unsigned)nseq, kHeads / kWarpsPerBlock), block, 0, st>>>(
        (const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (bf16*)out, ldo, Tq, Tk, mask_last);
  } else {
    NPVP_REQUIRE(false, "npvp_attention: unknown mode %d", mode);
  }
  NPVP_LAUNCH_CHECK("attention_kernel");
  return NPVP_OK;
}
