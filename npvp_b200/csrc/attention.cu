// Attention cores of the factorised VidHRFormer attention (8 heads x 64): 4x4 spatial windows (L = 16) and
// per-pixel temporal / encoder-decoder sequences (L = T <= 32).  Sequences are so short that the whole
// softmax(QK^T/8)V of one (sequence, head) lives in one warp: lane j owns key j (its K row in registers),
// scores are reduced with warp shuffles, P is staged in shared memory and lanes own output dims for PV.
// The window / time gathers are folded into the row index math, so no permuted copies are materialised.
// HBM-bound by design (reads Q,K,V once, writes O once); FLOPs are negligible next to the projections.
#include "common.cuh"

constexpr int kHeads = 8;
constexpr int kHeadsPerBlock = 4;   // 4 warps per block keeps static smem at 33 KB
constexpr int kHd = 64;
constexpr int kMaxL = 32;

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

template <int MODE>
__global__ void __launch_bounds__(128)
attention_kernel(const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k, int64_t ldk, const bf16* __restrict__ v,
                 int64_t ldv, bf16* __restrict__ out, int64_t ldo, int Tq, int Tk, int mask_last) {
  // per warp: Q [32][64] bf16, V [32][64] bf16, P [32] fp32
  __shared__ __align__(16) bf16 sQ[kHeadsPerBlock][kMaxL][kHd];
  __shared__ __align__(16) bf16 sV[kHeadsPerBlock][kMaxL][kHd];
  __shared__ float sP[kHeadsPerBlock][kMaxL];
  const int head = threadIdx.x >> 5, lane = threadIdx.x & 31;   // head = slot inside the block
  const int64_t s = blockIdx.x;
  const int Lq = (MODE == NPVP_ATTN_SPATIAL_WINDOW) ? 16 : Tq;
  const int Lk = (MODE == NPVP_ATTN_SPATIAL_WINDOW) ? 16 : Tk;
  const int col = (blockIdx.y * kHeadsPerBlock + head) * kHd;

  // stage Q and V rows of this head (128 B per row, one 4-byte word per lane)
  for (int i = 0; i < Lq; ++i) {
    const int64_t r = seq_row<MODE>(s, i, Tq);
    reinterpret_cast<uint32_t*>(&sQ[head][i][0])[lane] = __ldg(reinterpret_cast<const uint32_t*>(q + r * ldq + col) + lane);
  }
  for (int j = 0; j < Lk; ++j) {
    const int64_t r = seq_row<MODE>(s, j, Tk);
    reinterpret_cast<uint32_t*>(&sV[head][j][0])[lane] = __ldg(reinterpret_cast<const uint32_t*>(v + r * ldv + col) + lane);
  }
  // lane j keeps K_j (64 bf16) in 32 packed registers
  uint32_t kreg[32];
  if (lane < Lk) {
    const int64_t r = seq_row<MODE>(s, lane, Tk);
    const uint4* src = reinterpret_cast<const uint4*>(k + r * ldk + col);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const uint4 u = __ldg(src + t);
      kreg[4 * t] = u.x; kreg[4 * t + 1] = u.y; kreg[4 * t + 2] = u.z; kreg[4 * t + 3] = u.w;
    }
  } else {
#pragma unroll
    for (int t = 0; t < 32; ++t) kreg[t] = 0u;
  }
  __syncwarp();

  for (int i = 0; i < Lq; ++i) {
    // score_ij = (q_i / 8) . k_j
    float acc = 0.f;
    const uint4* qrow = reinterpret_cast<const uint4*>(&sQ[head][i][0]);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const uint4 u = qrow[t];   // broadcast read
      const uint32_t qw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = unpack_bf16x2(qw[e]);
        const float2 b = unpack_bf16x2(kreg[4 * t + e]);
        acc = fmaf(a.x, b.x, acc);
        acc = fmaf(a.y, b.y, acc);
      }
    }
    acc *= 0.125f;
    const bool masked = (lane >= Lk) || (mask_last && lane == Lk - 1 && i < Lq - 1);
    const float sc = masked ? -INFINITY : acc;
    const float mx = warp_max(sc);
    const float e = masked ? 0.f : expf(sc - mx);
    const float denom = warp_sum(e);
    sP[head][lane] = e / denom;
    __syncwarp();
    // o_i[d] = sum_j p_ij v_j[d]; lane owns d = 2*lane, 2*lane+1
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < Lk; ++j) {
      const float p = sP[head][j];
      const float2 vv = unpack_bf16x2(reinterpret_cast<const uint32_t*>(&sV[head][j][0])[lane]);
      o0 = fmaf(p, vv.x, o0);
      o1 = fmaf(p, vv.y, o1);
    }
    const int64_t r = seq_row<MODE>(s, i, Tq);
    reinterpret_cast<uint32_t*>(out + r * ldo + col)[lane] = pack_bf16x2(o0, o1);
    __syncwarp();
  }
}

extern "C" int npvp_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                              int64_t ldo, int mode, int64_t n_clips, int Tq, int Tk, int mask_last, void* stream) {
  NPVP_REQUIRE(q && k && v && out, "npvp_attention: null pointer");
  NPVP_REQUIRE(n_clips > 0 && Tq > 0 && Tk > 0, "npvp_attention: empty input");
  NPVP_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "npvp_attention: row strides must be multiples of 8");
  NPVP_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) % 16 == 0, "npvp_attention: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == NPVP_ATTN_SPATIAL_WINDOW) {
    NPVP_REQUIRE(Tq == Tk, "npvp_attention: spatial window attention needs Tq == Tk");
    const int64_t nseq = n_clips * Tq * 4;
    attention_kernel<NPVP_ATTN_SPATIAL_WINDOW><<<dim3((unsigned)nseq, kHeads / kHeadsPerBlock), 128, 0, st>>>((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv,
                                                                            (bf16*)out, ldo, Tq, Tk, 0);
  } else if (mode == NPVP_ATTN_TEMPORAL) {
    NPVP_REQUIRE(Tq <= kMaxL && Tk <= kMaxL, "npvp_attention: temporal length above %d not supported (Tq=%d Tk=%d)", kMaxL, Tq, Tk);
    const int64_t nseq = n_clips * 64;
    attention_kernel<NPVP_ATTN_TEMPORAL><<<dim3((unsigned)nseq, kHeads / kHeadsPerBlock), 128, 0, st>>>((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv,
                                                                      (bf16*)out, ldo, Tq, Tk, mask_last);
  } else {
    NPVP_REQUIRE(false, "npvp_attention: unknown mode %d", mode);
  }
  NPVP_LAUNCH_CHECK("attention_kernel");
  return NPVP_OK;
}
