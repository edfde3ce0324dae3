// Memory-bound kernels of the NP predictor: fused LayerNorm + positional GroupNorm fuse, per-frame
// LayerNorm(C,8,8)+GELU+residual, the conv-FFN middle (LN -> GELU -> depthwise 3x3 -> LN -> GELU),
// Fourier features, event-encoder pieces and layout changes.  Geometry is the reference's only one:
// 8x8 token grid (64 tokens per frame), embed dim C = 512.  All statistics are fp32 (double for the
// final mean/variance combine); bf16 is used only for buffers that feed tensor-core GEMMs.
#include "common.cuh"

constexpr int kC = 512;       // embed dim
constexpr int kTok = 64;      // tokens per frame (8x8)
constexpr float kEps = 1e-5f;

// Per-frame kernels (statistics over a whole 64 x 512 frame).  The first version held a frame in the registers of ONE
// 512-thread block (64 values per thread, 110-128 registers): one block per SM, so the load -> reduce -> compute -> store
// chain of a frame never overlapped with anything (24% warps active, 3.5 TB/s, r01 ncu).  A frame is now owned by a
// CLUSTER of kFrameCL blocks of 256 threads: block `rank` holds tokens 16 rank .. 16 rank + 15 (32 values per thread), the
// whole-frame mean / variance is merged across the cluster through distributed shared memory (st.async + mbarrier), and four such blocks of
// different frames share an SM.
// Thread layout: warp w of block `rank` owns tokens tok0 = 16 rank + 2 w and tok0 + 1; lane l owns channels
// {128 j + 4 l + i}: one float4 per j => fully coalesced 512-byte warp accesses.
constexpr int kFrameCL = 4;                                   // blocks per frame (cluster size)
constexpr int kTPW = 2;                                       // tokens per warp
constexpr int kFrameThreads = kTok / kFrameCL / kTPW * 32;    // 256
struct FrameRegs { float v[kTPW][16]; };

__device__ __forceinline__ uint32_t frame_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void frame_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t frame_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t frame_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

// Cluster exchange of the per-block partial statistics.  Each block owns, per exchange e, a table slots[e][kFrameCL] and an
// mbarrier bars[e].  Thread 0 posts its (mean, M2) into slot [e][own rank] of every PEER with st.async (a remote
// shared-memory store that completes transaction bytes on the peer's mbarrier), writes its own entry locally and arrives
// with expect_tx; every thread then waits on the LOCAL mbarrier and reads the LOCAL table.  Unlike barrier.cluster with
// release / acquire semantics this does not make every thread drain its outstanding global stores first (r01 ncu: the
// second exchange of the TAIL kernel sat behind the y / out_ln stores: stall_membar 1.8-3.8 and stall_barrier 1.6-3.5
// issue slots per instruction).  The only cluster barrier left is the one that publishes the mbarrier initialisation at
// kernel entry, when nothing is in flight.  No block can exit while a peer still writes into it: it has waited for all of
// them.
struct FrameXchg {
  float2 slots[2][kFrameCL];
  unsigned long long bars[2];
};

__device__ __forceinline__ void frame_xchg_init(FrameXchg& x) {
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(frame_smem_u32(&x.bars[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(frame_smem_u32(&x.bars[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  frame_cluster_sync();
}

// thread 0 of every block calls post(); all threads call wait() and then read x.slots[e][0..kFrameCL)
__device__ __forceinline__ void frame_xchg_post(FrameXchg& x, int e, float2 v) {
  const uint32_t me = frame_cluster_rank();
  const uint32_t bar = frame_smem_u32(&x.bars[e]), slot = frame_smem_u32(&x.slots[e][me]);
  x.slots[e][me] = v;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)((kFrameCL - 1) * sizeof(float2))) : "memory");
#pragma unroll
  for (uint32_t r = 0; r < (uint32_t)kFrameCL; ++r) {
    if (r == me) continue;
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(frame_mapa(slot, r)), "f"(v.x),
                 "f"(v.y), "r"(frame_mapa(bar, r))
                 : "memory");
  }
}
__device__ __forceinline__ void frame_xchg_wait(FrameXchg& x, int e) {
  const uint32_t bar = frame_smem_u32(&x.bars[e]);
  uint32_t ok = 0, spins = 0;
  long long t0 = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar)
        : "memory");
    if (!ok && (++spins & 0x3FFu) == 0) {                      // bounded: a lost arrival traps instead of hanging the GPU
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) __trap();
    }
  }
}

// x: first token of this warp (row pointer arithmetic is done by the callers through tok0)
__device__ __forceinline__ void frame_load(const float* __restrict__ x, FrameRegs& r, int tok0, int lane) {
#pragma unroll
  for (int t = 0; t < kTPW; ++t) {
    const float4* row = reinterpret_cast<const float4*>(x + (size_t)(tok0 + t) * kC);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 q = __ldg(row + j * 32 + lane);
      r.v[t][4 * j + 0] = q.x; r.v[t][4 * j + 1] = q.y; r.v[t][4 * j + 2] = q.z; r.v[t][4 * j + 3] = q.w;
    }
  }
}

__device__ __forceinline__ void frame_load_bf16(const bf16* __restrict__ x, FrameRegs& r, int tok0, int lane) {
#pragma unroll
  for (int t = 0; t < kTPW; ++t) {
    const uint2* row = reinterpret_cast<const uint2*>(x + (size_t)(tok0 + t) * kC);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint2 u = __ldg(row + j * 32 + lane);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
      r.v[t][4 * j + 0] = a.x; r.v[t][4 * j + 1] = a.y; r.v[t][4 * j + 2] = b.x; r.v[t][4 * j + 3] = b.y;
    }
  }
}

__device__ __forceinline__ void frame_store_bf16(bf16* __restrict__ out, const FrameRegs& r, int tok0, int lane) {
#pragma unroll
  for (int t = 0; t < kTPW; ++t) {
    uint2* row = reinterpret_cast<uint2*>(out + (size_t)(tok0 + t) * kC);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      row[j * 32 + lane] = make_uint2(pack_bf16x2(r.v[t][4 * j], r.v[t][4 * j + 1]), pack_bf16x2(r.v[t][4 * j + 2], r.v[t][4 * j + 3]));
  }
}

// per-token LayerNorm over 512 channels held as 16 values per lane
__device__ __forceinline__ void token_layernorm(float (&v)[16], const float4 (&w)[4], const float4 (&b)[4]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  const float mu = warp_sum(s) * (1.0f / kC);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) { const float d = v[i] - mu; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / kC) + kEps);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[4 * j + 0] = (v[4 * j + 0] - mu) * rstd * w[j].x + b[j].x;
    v[4 * j + 1] = (v[4 * j + 1] - mu) * rstd * w[j].y + b[j].y;
    v[4 * j + 2] = (v[4 * j + 2] - mu) * rstd * w[j].z + b[j].z;
    v[4 * j + 3] = (v[4 * j + 3] - mu) * rstd * w[j].w + b[j].w;
  }
}

// Mean / rstd over the whole frame.  Each block reduces its quarter with a two-pass (mean, centred sum of squares),
// exchanges (mean_k, M2_k) with its three peers (frame_xchg_*), and every thread merges the four quarters
// (equal counts: M2 = sum M2_k + n_k sum (mean_k - mean)^2).
__device__ __forceinline__ void frame_stats(const FrameRegs& r, float* red, FrameXchg& xc, int e, float& mean, float& rstd) {
  constexpr float inv_nk = 1.0f / (kTok / kFrameCL * kC);
  float s = 0.f, dummy = 0.f;
#pragma unroll
  for (int t = 0; t < kTPW; ++t)
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r.v[t][i];
  block_sum2(s, dummy, red);
  const float mk = s * inv_nk;
  float q = 0.f;
  dummy = 0.f;
#pragma unroll
  for (int t = 0; t < kTPW; ++t)
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float d = r.v[t][i] - mk; q = fmaf(d, d, q); }
  block_sum2(q, dummy, red);
  if (threadIdx.x == 0) frame_xchg_post(xc, e, make_float2(mk, q));
  frame_xchg_wait(xc, e);
  float2 part[kFrameCL];
#pragma unroll
  for (int k = 0; k < kFrameCL; ++k) part[k] = xc.slots[e][k];
  float m = 0.f;
#pragma unroll
  for (int k = 0; k < kFrameCL; ++k) m += part[k].x;
  mean = m * (1.0f / kFrameCL);
  float m2 = 0.f, dm = 0.f;
#pragma unroll
  for (int k = 0; k < kFrameCL; ++k) { m2 += part[k].y; const float d = part[k].x - mean; dm = fmaf(d, d, dm); }
  rstd = rsqrtf(m2 * (1.0f / (kTok * kC)) + dm * (1.0f / kFrameCL) + kEps);
}

// ---------------------------------------------------------------------------------------------
// a = LN(x); u = a + qe; fused = GN1(u) * (1 + gamma) + beta
// ---------------------------------------------------------------------------------------------
// r holds this block's quarter of the frame in fp32; applies the optional token LayerNorm, writes `a`, then the positional fuse.
__device__ __forceinline__ void posfuse_from_regs(FrameRegs& r, float* red, FrameXchg& xc, int f, int T, int Tpos, const float* __restrict__ ln_w,
                                                  const float* __restrict__ ln_b, const float* __restrict__ qe,
                                                  const float* __restrict__ beta, const float* __restrict__ gamma,
                                                  bf16* __restrict__ out_ln, bf16* __restrict__ out_fused, int tok0, int lane) {
  const int n = f / T, t_idx = f % Tpos;      // Tpos = T: one timestamp set for the batch; n_clips * T: per-clip timestamps
  if (ln_w) {
    float4 w[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      w[j] = __ldg(reinterpret_cast<const float4*>(ln_w) + j * 32 + lane);
      b[j] = __ldg(reinterpret_cast<const float4*>(ln_b) + j * 32 + lane);
    }
#pragma unroll
    for (int t = 0; t < kTPW; ++t) token_layernorm(r.v[t], w, b);
  }
  if (out_ln) frame_store_bf16(out_ln + (size_t)f * kTok * kC, r, tok0, lane);
  if (!out_fused) return;
  if (qe) {
#pragma unroll
    for (int t = 0; t < kTPW; ++t) {
      const float4* row = reinterpret_cast<const float4*>(qe + ((size_t)n * kTok + tok0 + t) * kC);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 q = __ldg(row + j * 32 + lane);
        r.v[t][4 * j] += q.x; r.v[t][4 * j + 1] += q.y; r.v[t][4 * j + 2] += q.z; r.v[t][4 * j + 3] += q.w;
      }
    }
  }
  float mean, rstd;
  frame_stats(r, red, xc, 1, mean, rstd);
#pragma unroll
  for (int t = 0; t < kTPW; ++t) {
    const size_t off = ((size_t)t_idx * kTok + tok0 + t) * kC;
    const float4* brow = reinterpret_cast<const float4*>(beta + off);
    const float4* grow = gamma ? reinterpret_cast<const float4*>(gamma + off) : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 be = __ldg(brow + j * 32 + lane);
      float4 ga = make_float4(0.f, 0.f, 0.f, 0.f);
      if (grow) ga = __ldg(grow + j * 32 + lane);
      r.v[t][4 * j + 0] = (r.v[t][4 * j + 0] - mean) * rstd * (1.0f + ga.x) + be.x;
      r.v[t][4 * j + 1] = (r.v[t][4 * j + 1] - mean) * rstd * (1.0f + ga.y) + be.y;
      r.v[t][4 * j + 2] = (r.v[t][4 * j + 2] - mean) * rstd * (1.0f + ga.z) + be.z;
      r.v[t][4 * j + 3] = (r.v[t][4 * j + 3] - mean) * rstd * (1.0f + ga.w) + be.w;
    }
  }
  frame_store_bf16(out_fused + (size_t)f * kTok * kC, r, tok0, lane);
}

// x += delta (a residual branch's output left in bf16 by its GEMM), written back in fp32, for the quarter held in registers
__device__ __forceinline__ void frame_add_delta(float* __restrict__ x, const bf16* __restrict__ delta, FrameRegs& r, int tok0, int lane) {
#pragma unroll
  for (int t = 0; t < kTPW; ++t) {
    const uint2* drow = reinterpret_cast<const uint2*>(delta + (size_t)(tok0 + t) * kC);
    float4* xrow = reinterpret_cast<float4*>(x + (size_t)(tok0 + t) * kC);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint2 u = __ldg(drow + j * 32 + lane);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
      r.v[t][4 * j] += a.x; r.v[t][4 * j + 1] += a.y; r.v[t][4 * j + 2] += b.x; r.v[t][4 * j + 3] += b.y;
      xrow[j * 32 + lane] = make_float4(r.v[t][4 * j], r.v[t][4 * j + 1], r.v[t][4 * j + 2], r.v[t][4 * j + 3]);
    }
  }
}

__global__ void __cluster_dims__(kFrameCL, 1, 1) __launch_bounds__(kFrameThreads, 4)
ln_posfuse_kernel(float* __restrict__ x, const bf16* __restrict__ delta, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                  const float* __restrict__ qe, const float* __restrict__ beta, const float* __restrict__ gamma,
                  bf16* __restrict__ out_ln, bf16* __restrict__ out_fused, int T, int Tpos) {
  __shared__ float red[64];
  __shared__ __align__(8) FrameXchg xc;
  frame_xchg_init(xc);
  const int f = blockIdx.x / kFrameCL, lane = threadIdx.x & 31;
  const int tok0 = (int)frame_cluster_rank() * (kTok / kFrameCL) + (threadIdx.x >> 5) * kTPW;
  FrameRegs r;
  frame_load(x + (size_t)f * kTok * kC, r, tok0, lane);
  if (delta) frame_add_delta(x + (size_t)f * kTok * kC, delta + (size_t)f * kTok * kC, r, tok0, lane);
  posfuse_from_regs(r, red, xc, f, T, Tpos, ln_w, ln_b, qe, beta, gamma, out_ln, out_fused, tok0, lane);
}

extern "C" int npvp_ln_posfuse(const float* x, const float* ln_w, const float* ln_b, const float* qe, const float* beta,
                               const float* gamma, void* out_ln_bf16, void* out_fused_bf16, int64_t n_clips, int64_t T,
                               int64_t pos_frames, void* stream) {
  NPVP_REQUIRE(x && (out_ln_bf16 || out_fused_bf16), "npvp_ln_posfuse: null pointer");
  NPVP_REQUIRE((ln_w == nullptr) == (ln_b == nullptr), "npvp_ln_posfuse: ln_w/ln_b must both be set or both NULL");
  NPVP_REQUIRE(!out_fused_bf16 || beta, "npvp_ln_posfuse: beta required for the fused output");
  NPVP_REQUIRE(n_clips > 0 && T > 0, "npvp_ln_posfuse: empty input");
  NPVP_REQUIRE(pos_frames == 0 || pos_frames == T || pos_frames == n_clips * T, "npvp_ln_posfuse: pos_frames must be 0, T or n_clips * T");
  ln_posfuse_kernel<<<(unsigned)(n_clips * T * kFrameCL), kFrameThreads, 0, (cudaStream_t)stream>>>(const_cast<float*>(x), nullptr, ln_w, ln_b, qe, beta, gamma,
                                                                              (bf16*)out_ln_bf16, (bf16*)out_fused_bf16, (int)T, (int)(pos_frames ? pos_frames : T));
  NPVP_LAUNCH_CHECK("ln_posfuse_kernel");
  return NPVP_OK;
}

extern "C" int npvp_add_ln_posfuse(float* x, const void* delta_bf16, const float* ln_w, const float* ln_b, const float* qe,
                                   const float* beta, const float* gamma, void* out_ln_bf16, void* out_fused_bf16, int64_t n_clips,
                                   int64_t T, int64_t pos_frames, void* stream) {
  NPVP_REQUIRE(x && delta_bf16 && (out_ln_bf16 || out_fused_bf16), "npvp_add_ln_posfuse: null pointer");
  NPVP_REQUIRE((ln_w == nullptr) == (ln_b == nullptr), "npvp_add_ln_posfuse: ln_w/ln_b must both be set or both NULL");
  NPVP_REQUIRE(!out_fused_bf16 || beta, "npvp_add_ln_posfuse: beta required for the fused output");
  NPVP_REQUIRE(n_clips > 0 && T > 0, "npvp_add_ln_posfuse: empty input");
  NPVP_REQUIRE(pos_frames == 0 || pos_frames == T || pos_frames == n_clips * T, "npvp_add_ln_posfuse: pos_frames must be 0, T or n_clips * T");
  ln_posfuse_kernel<<<(unsigned)(n_clips * T * kFrameCL), kFrameThreads, 0, (cudaStream_t)stream>>>(x, (const bf16*)delta_bf16, ln_w, ln_b, qe, beta, gamma,
                                                                              (bf16*)out_ln_bf16, (bf16*)out_fused_bf16, (int)T, (int)(pos_frames ? pos_frames : T));
  NPVP_LAUNCH_CHECK("ln_posfuse_kernel<add>");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm(512) per token: one warp per row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(float* __restrict__ x, const bf16* __restrict__ delta, const float* __restrict__ w, const float* __restrict__ b,
                      float* __restrict__ out_f32, h16* __restrict__ out_bf16, int64_t rows, int relu, int fp16) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[16];
  float4 wv[4], bv[4];
  float4* src = reinterpret_cast<float4*>(x + row * kC);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float4 q = src[j * 32 + lane];
    if (delta) {                                   // deferred residual add: x += delta, written back
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(delta + row * kC) + j * 32 + lane);
      const float2 da = unpack_bf16x2(u.x), db = unpack_bf16x2(u.y);
      q.x += da.x; q.y += da.y; q.z += db.x; q.w += db.y;
      src[j * 32 + lane] = q;
    }
    v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
    wv[j] = __ldg(reinterpret_cast<const float4*>(w) + j * 32 + lane);
    bv[j] = __ldg(reinterpret_cast<const float4*>(b) + j * 32 + lane);
  }
  token_layernorm(v, wv, bv);
  if (relu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (out_f32) reinterpret_cast<float4*>(out_f32 + row * kC)[j * 32 + lane] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    if (out_bf16) reinterpret_cast<uint2*>(out_bf16 + row * kC)[j * 32 + lane] = make_uint2(pack_h16x2(v[4 * j], v[4 * j + 1], fp16), pack_h16x2(v[4 * j + 2], v[4 * j + 3], fp16));
  }
}

extern "C" int npvp_layernorm_rows(const float* x, const float* w, const float* b, float* out_f32, void* out_bf16,
                                   int64_t rows, int relu, int fp16, void* stream) {
  NPVP_REQUIRE(x && w && b && (out_f32 || out_bf16) && rows > 0, "npvp_layernorm_rows: bad arguments");
  layernorm_rows_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, (cudaStream_t)stream>>>(const_cast<float*>(x), nullptr, w, b, out_f32, (h16*)out_bf16, rows, relu, fp16);
  NPVP_LAUNCH_CHECK("layernorm_rows_kernel");
  return NPVP_OK;
}

extern "C" int npvp_add_layernorm_rows(float* x, const void* delta_bf16, const float* w, const float* b, float* out_f32, void* out_bf16,
                                       int64_t rows, int relu, int fp16, void* stream) {
  NPVP_REQUIRE(x && delta_bf16 && w && b && (out_f32 || out_bf16) && rows > 0, "npvp_add_layernorm_rows: bad arguments");
  NPVP_REQUIRE(out_f32 != x, "npvp_add_layernorm_rows: out_f32 must not alias x (x receives x + delta)");
  layernorm_rows_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, (const bf16*)delta_bf16, w, b, out_f32, (h16*)out_bf16, rows, relu, fp16);
  NPVP_LAUNCH_CHECK("layernorm_rows_kernel<add>");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// y += GELU(LayerNorm_(C,8,8)(h) * w + b)
// ---------------------------------------------------------------------------------------------
template <bool TAIL>
__global__ void __cluster_dims__(kFrameCL, 1, 1) __launch_bounds__(kFrameThreads, 4)
frame_ln_gelu_residual_kernel(const void* __restrict__ h, int h_is_bf16, const float* __restrict__ w_hwc, const float* __restrict__ b_hwc,
                              float* __restrict__ y, int T, int Tpos, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                              const float* __restrict__ qe, const float* __restrict__ beta, const float* __restrict__ gamma,
                              bf16* __restrict__ out_ln, bf16* __restrict__ out_fused) {
  __shared__ float red[64];
  __shared__ __align__(8) FrameXchg xc;
  frame_xchg_init(xc);
  const int f = blockIdx.x / kFrameCL, lane = threadIdx.x & 31;
  const int tok0 = (int)frame_cluster_rank() * (kTok / kFrameCL) + (threadIdx.x >> 5) * kTPW;
  FrameRegs r;
  if (h_is_bf16) frame_load_bf16((const bf16*)h + (size_t)f * kTok * kC, r, tok0, lane);
  else frame_load((const float*)h + (size_t)f * kTok * kC, r, tok0, lane);
  float mean, rstd;
  frame_stats(r, red, xc, 0, mean, rstd);
#pragma unroll
  for (int t = 0; t < kTPW; ++t) {
    const size_t tok = tok0 + t;
    const float4* wrow = reinterpret_cast<const float4*>(w_hwc + tok * kC);
    const float4* brow = reinterpret_cast<const float4*>(b_hwc + tok * kC);
    float4* yrow = reinterpret_cast<float4*>(y + ((size_t)f * kTok + tok) * kC);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 ww = __ldg(wrow + j * 32 + lane), bb = __ldg(brow + j * 32 + lane);
      float4 yy = yrow[j * 32 + lane];
      yy.x += gelu_erf((r.v[t][4 * j + 0] - mean) * rstd * ww.x + bb.x);
      yy.y += gelu_erf((r.v[t][4 * j + 1] - mean) * rstd * ww.y + bb.y);
      yy.z += gelu_erf((r.v[t][4 * j + 2] - mean) * rstd * ww.z + bb.z);
      yy.w += gelu_erf((r.v[t][4 * j + 3] - mean) * rstd * ww.w + bb.w);
      yrow[j * 32 + lane] = yy;
      if (TAIL) { r.v[t][4 * j] = yy.x; r.v[t][4 * j + 1] = yy.y; r.v[t][4 * j + 2] = yy.z; r.v[t][4 * j + 3] = yy.w; }
    }
  }
  // fused consumer: the next op of every block is LayerNorm + positional fuse of the stream just updated
  if (TAIL) posfuse_from_regs(r, red, xc, f, T, Tpos, ln_w, ln_b, qe, beta, gamma, out_ln, out_fused, tok0, lane);
}

extern "C" int npvp_frame_ln_gelu_residual(const void* h, int h_is_bf16, const float* w_hwc, const float* b_hwc, float* y, int64_t frames,
                                           void* stream) {
  NPVP_REQUIRE(h && w_hwc && b_hwc && y && frames > 0, "npvp_frame_ln_gelu_residual: bad arguments");
  frame_ln_gelu_residual_kernel<false><<<(unsigned)(frames * kFrameCL), kFrameThreads, 0, (cudaStream_t)stream>>>(h, h_is_bf16, w_hwc, b_hwc, y, 1, 1, nullptr, nullptr, nullptr,
                                                                                          nullptr, nullptr, nullptr, nullptr);
  NPVP_LAUNCH_CHECK("frame_ln_gelu_residual_kernel");
  return NPVP_OK;
}

extern "C" int npvp_frame_ln_gelu_residual_posfuse(const void* h, int h_is_bf16, const float* w_hwc, const float* b_hwc, float* y, const float* ln_w,
                                                   const float* ln_b, const float* qe, const float* beta, const float* gamma,
                                                   void* out_ln_bf16, void* out_fused_bf16, int64_t n_clips, int64_t T, int64_t pos_frames,
                                                   void* stream) {
  NPVP_REQUIRE(h && w_hwc && b_hwc && y && n_clips > 0 && T > 0, "npvp_frame_ln_gelu_residual_posfuse: bad arguments");
  NPVP_REQUIRE((ln_w == nullptr) == (ln_b == nullptr), "npvp_frame_ln_gelu_residual_posfuse: ln_w/ln_b must both be set or both NULL");
  NPVP_REQUIRE((out_ln_bf16 || out_fused_bf16) && (!out_fused_bf16 || beta), "npvp_frame_ln_gelu_residual_posfuse: outputs / beta missing");
  NPVP_REQUIRE(pos_frames == 0 || pos_frames == T || pos_frames == n_clips * T,
               "npvp_frame_ln_gelu_residual_posfuse: pos_frames must be 0, T or n_clips * T");
  frame_ln_gelu_residual_kernel<true><<<(unsigned)(n_clips * T * kFrameCL), kFrameThreads, 0, (cudaStream_t)stream>>>(
      h, h_is_bf16, w_hwc, b_hwc, y, (int)T, (int)(pos_frames ? pos_frames : T), ln_w, ln_b, qe, beta, gamma, (bf16*)out_ln_bf16,
      (bf16*)out_fused_bf16);
  NPVP_LAUNCH_CHECK("frame_ln_gelu_residual_kernel<posfuse>");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// temporal mean of the memory
// ---------------------------------------------------------------------------------------------
__global__ void temporal_mean_kernel(const float4* __restrict__ mem, float4* __restrict__ evt, int64_t n_clips, int T, int64_t fe4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_clips * fe4) return;
  const int64_t n = i / fe4, e = i % fe4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < T; ++t) {
    const float4 q = __ldg(mem + (n * T + t) * fe4 + e);
    s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
  }
  const float inv = 1.0f / (float)T;
  evt[i] = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
}

extern "C" int npvp_temporal_mean(const float* mem, float* evt, int64_t n_clips, int64_t T, int64_t frame_elems, void* stream) {
  NPVP_REQUIRE(mem && evt && n_clips > 0 && T > 0 && frame_elems % 4 == 0, "npvp_temporal_mean: bad arguments");
  const int64_t total = n_clips * (frame_elems / 4);
  temporal_mean_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)mem, (float4*)evt, n_clips, (int)T, frame_elems / 4);
  NPVP_LAUNCH_CHECK("temporal_mean_kernel");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// conv-FFN middle
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ffn_frame_stats_kernel(const bf16* __restrict__ h, float* __restrict__ stats, int64_t elems) {
  __shared__ float red[64];
  const uint4* src = reinterpret_cast<const uint4*>(h + (size_t)blockIdx.x * elems);
  float s = 0.f, q = 0.f;
  for (int64_t i = threadIdx.x; i < elems / 8; i += 256) {
    const uint4 u = __ldg(src + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 p = unpack_bf16x2(w[k]);
      s += p.x + p.y;
      q = fmaf(p.x, p.x, q);
      q = fmaf(p.y, p.y, q);
    }
  }
  block_sum2(s, q, red);
  if (threadIdx.x == 0) {
    const double mean = (double)s / (double)elems;
    const double var = fmax((double)q / (double)elems - mean * mean, 0.0);
    stats[2 * blockIdx.x] = (float)mean;
    stats[2 * blockIdx.x + 1] = (float)(1.0 / sqrt(var + (double)kEps));
  }
}

extern "C" int npvp_ffn_frame_stats(const void* h_bf16, float* stats, int64_t frames, int64_t Ch, void* stream) {
  NPVP_REQUIRE(h_bf16 && stats && frames > 0 && Ch % 8 == 0, "npvp_ffn_frame_stats: bad arguments");
  ffn_frame_stats_kernel<<<(unsigned)frames, 256, 0, (cudaStream_t)stream>>>((const bf16*)h_bf16, stats, kTok * Ch);
  NPVP_LAUNCH_CHECK("ffn_frame_stats_kernel");
  return NPVP_OK;
}

// (mean, rstd) of a frame from the partial sums a GEMM epilogue left behind (npvp_epilogue_t.frame_stats): one thread per frame
__global__ void ffn_stats_finalize_kernel(const float2* __restrict__ partial, int P, float* __restrict__ stats, int64_t frames, double inv_n) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= frames) return;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < P; ++k) {
    const float2 v = __ldg(partial + f * P + k);
    s += (double)v.x;
    q += (double)v.y;
  }
  const double mean = s * inv_n;
  stats[2 * f] = (float)mean;
  stats[2 * f + 1] = (float)(1.0 / sqrt(fmax(q * inv_n - mean * mean, 0.0) + (double)kEps));
}

extern "C" int npvp_ffn_stats_finalize(const float* partial, int64_t P, float* stats, int64_t frames, int64_t elems_per_frame, void* stream) {
  NPVP_REQUIRE(partial && stats && P > 0 && frames > 0 && elems_per_frame > 0, "npvp_ffn_stats_finalize: bad arguments");
  ffn_stats_finalize_kernel<<<(unsigned)ceil_div64(frames, 128), 128, 0, (cudaStream_t)stream>>>((const float2*)partial, (int)P, stats, frames,
                                                                                                 1.0 / (double)elems_per_frame);
  NPVP_LAUNCH_CHECK("ffn_stats_finalize_kernel");
  return NPVP_OK;
}

// depthwise 3x3, zero padding 1, on an 8x8 map held in registers (a[p], p = y*8+x); taps w[ky*3+kx]
__device__ __forceinline__ float dw3x3_at(const float (&a)[64], const float (&w)[9], float bias, int y, int x) {
  float acc = bias;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = y + ky - 1;
    if (yy < 0 || yy > 7) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = x + kx - 1;
      if (xx < 0 || xx > 7) continue;
      acc = fmaf(a[yy * 8 + xx], w[ky * 3 + kx], acc);
    }
  }
  return acc;
}

// One thread per channel: normalise + GELU the 8x8 map, depthwise conv, write bf16, per-block partial stats.
// LayerNorm((Ch,8,8)) has a weight and a bias PER ELEMENT of the frame, i.e. 8 B of fp32 parameters for every 2 B activation;
// read per frame from L2 they, not HBM, bounded the first version (146 us per 640 frames, 2.8x the HBM time).  Here a block
// owns kFfnChunk channels, parks their 64 x 2 parameters in shared memory once (each thread only ever reads its own
// column, so no barrier is needed) and walks over frames blockIdx.y, blockIdx.y + gridDim.y, ...
constexpr int kFfnChunk = 128;                                    // channels per block
constexpr int kFfnGroups = 5;                                     // frames in flight per block (128 threads each)
template <int CH>                                                 // CH > 0: compile-time channel count (row stride becomes an immediate)
__global__ void __launch_bounds__(kFfnChunk * kFfnGroups, 1)
ffn_dwconv_kernel(const bf16* __restrict__ h, const float* __restrict__ stats1, const float* __restrict__ n1w,
                  const float* __restrict__ n1b, const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                  bf16* __restrict__ y, float* __restrict__ partial2, int Ch_rt, int frames) {
  // The 64 KB of LayerNorm parameters of a 128-channel chunk are shared by kFfnGroups groups of 128 threads, each walking
  // over its own frames: 20 warps per SM instead of the 12 that three independent 128-thread blocks (3 x 64 KB) allowed -
  // the kernel is bound by per-warp instruction latency (r01 ncu: 2.9 warps per scheduler, 52% issue slots used).
  extern __shared__ float2 ffn_wb[];                             // [64 px][kFfnChunk] (weight, bias)
  __shared__ float red[2][kFfnGroups][8];
  const int Ch = CH > 0 ? CH : Ch_rt;
  const int grp = threadIdx.x / kFfnChunk, tc = threadIdx.x % kFfnChunk, gw = tc >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * kFfnChunk + tc;
  for (int i = threadIdx.x; i < kTok * kFfnChunk; i += kFfnChunk * kFfnGroups) {
    const int p = i / kFfnChunk, cc = i % kFfnChunk;
    const size_t off = (size_t)p * Ch + blockIdx.x * kFfnChunk + cc;
    ffn_wb[i] = make_float2(__ldg(n1w + off), __ldg(n1b + off));
  }
  float w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w[k] = __ldg(dw_w + (size_t)k * Ch + c);
  const float bias = __ldg(dw_b + c);
  __syncthreads();
  int it = 0;
  for (int f = blockIdx.y * kFfnGroups + grp; f < frames; f += gridDim.y * kFfnGroups, it ^= 1) {
    const float rstd = __ldg(stats1 + 2 * f + 1), nmr = -__ldg(stats1 + 2 * f) * rstd;
    float a[64];
    const bf16* src = h + (size_t)f * kTok * Ch + c;
#pragma unroll
    for (int p = 0; p < 64; ++p) a[p] = __bfloat162float(src[(size_t)p * Ch]);
#pragma unroll
    for (int r = 0; r < 8; ++r) {                                   // one image row (8 independent chains) at a time
      float2 wb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) wb[j] = ffn_wb[(r * 8 + j) * kFfnChunk + tc];
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(fmaf(a[r * 8 + j], rstd, nmr), wb[j].x, wb[j].y);   // LN1 affine in two FMAs
      gelu_erf_x8(v);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[r * 8 + j] = v[j];
    }
    bf16* dst = y + (size_t)f * kTok * Ch + c;
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int yy = 0; yy < 8; ++yy) {
      float o[8];
#pragma unroll
      for (int xx = 0; xx < 8; ++xx) o[xx] = bias;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = yy + ky - 1;
        if (iy < 0 || iy > 7) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int xx = 0; xx < 8; ++xx) {
            const int ix = xx + kx - 1;
            if (ix >= 0 && ix <= 7) o[xx] = fmaf(a[iy * 8 + ix], w[ky * 3 + kx], o[xx]);
          }
      }
#pragma unroll
      for (int xx = 0; xx < 8; ++xx) {
        dst[(size_t)(yy * 8 + xx) * Ch] = __float2bfloat16(o[xx]);
        s += o[xx];
        q = fmaf(o[xx], o[xx], q);
      }
    }
    // per-group reduction (the groups run different frames, so no block-wide barrier): named barrier 1 + grp, buffers
    // alternate so a warp one iteration ahead cannot overwrite what thread 0 of the group is still reading
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane == 0) { red[it][grp][gw] = s; red[it][grp][4 + gw] = q; }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(kFfnChunk) : "memory");
    if (tc == 0) {
      float* p = partial2 + ((size_t)f * gridDim.x + blockIdx.x) * 2;
      p[0] = (red[it][grp][0] + red[it][grp][1]) + (red[it][grp][2] + red[it][grp][3]);
      p[1] = (red[it][grp][4] + red[it][grp][5]) + (red[it][grp][6] + red[it][grp][7]);
    }
  }
}

// Packed (FFMA2) version of ffn_dwconv_kernel, Ch = 2048.  The scalar kernel above issues ~27 instructions per element of
// which ~19 are FMA-pipe, and sits at the issue limit of its 20 warps, not at the HBM limit (97 us vs 52 us per 640 frames).
// Here a thread owns a channel PAIR and half the image, so every LayerNorm / GELU / tap / statistics step is one packed
// f32x2 instruction for two values and the loads / stores are 4-byte bf16x2: lane & 15 = channel pair, lane >> 4 = image
// half.  The lower half keeps its rows mirrored (local row r = image row 7 - r) and loads its vertical taps mirrored, so
// for both halves local row 0 touches the zero padding and local row 3 needs the other half's local row 3: one
// __shfl_xor(.., 16) per value, no divergence.  Output rows are produced bottom-up so the exchanged row dies first.
constexpr int kDw2Groups = 4;
constexpr int kDw2Pairs = kFfnChunk / 2;
__global__ void __launch_bounds__(kFfnChunk * kDw2Groups, 1)
ffn_dwconv2_kernel(const bf16* __restrict__ h, const float* __restrict__ stats1, const float* __restrict__ n1w,
                   const float* __restrict__ n1b, const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                   bf16* __restrict__ y, float* __restrict__ partial2, int frames) {
  constexpr int Ch = 2048;
  extern __shared__ float4 dw2_wb[];                             // [64 px][64 pairs] (w_c, w_c+1, b_c, b_c+1)
  __shared__ float red[2][kDw2Groups][8];
  const int grp = threadIdx.x / kFfnChunk, tc = threadIdx.x % kFfnChunk, gw = tc >> 5, lane = threadIdx.x & 31;
  const int pair = gw * 16 + (lane & 15), half = lane >> 4;
  const int c0 = blockIdx.x * kFfnChunk + 2 * pair;
  const int pbase = half ? 56 : 0, pstep = half ? -8 : 8;          // local row r = image row (half ? 7 - r : r)
  for (int i = threadIdx.x; i < kTok * kDw2Pairs; i += kFfnChunk * kDw2Groups) {
    const int p = i / kDw2Pairs, pr = i % kDw2Pairs;
    const size_t off = (size_t)p * Ch + blockIdx.x * kFfnChunk + 2 * pr;
    const float2 w = __ldg(reinterpret_cast<const float2*>(n1w + off));
    const float2 b = __ldg(reinterpret_cast<const float2*>(n1b + off));
    dw2_wb[i] = make_float4(w.x, w.y, b.x, b.y);
  }
  f32x2 w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int kk = (half ? 2 - k / 3 : k / 3) * 3 + k % 3;          // lower half: vertical taps mirrored
    const float2 t = __ldg(reinterpret_cast<const float2*>(dw_w + (size_t)kk * Ch + c0));
    w[k] = pk2(t.x, t.y);
  }
  const float2 bias_f = __ldg(reinterpret_cast<const float2*>(dw_b + c0));
  const f32x2 bias = pk2(bias_f.x, bias_f.y);
  __syncthreads();
  int it = 0;
  for (int f = blockIdx.y * kDw2Groups + grp; f < frames; f += gridDim.y * kDw2Groups, it ^= 1) {
    const float rstd1 = __ldg(stats1 + 2 * f + 1), nmr1 = -__ldg(stats1 + 2 * f) * rstd1;
    const f32x2 rs1 = pk2(rstd1, rstd1), nm1 = pk2(nmr1, nmr1);
    const bf16* src = h + (size_t)f * kTok * Ch + c0;
    uint32_t raw[32];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int x = 0; x < 8; ++x) raw[r * 8 + x] = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)(pbase + r * pstep + x) * Ch));
    f32x2 a[5][8];                                               // local rows 0..3 + the other half's row 3
#pragma unroll
    for (int r = 3; r >= 0; --r)
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        const float4 wb = dw2_wb[(pbase + r * pstep + x) * kDw2Pairs + pair];
        a[r][x] = gelu_erf2(fma2(fma2(bf16x2_to_f32x2(raw[r * 8 + x]), rs1, nm1), pk2(wb.x, wb.y), pk2(wb.z, wb.w)));
      }
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      float lo, hi;
      upk2(a[3][x], lo, hi);
      a[4][x] = pk2(__shfl_xor_sync(0xffffffffu, lo, 16), __shfl_xor_sync(0xffffffffu, hi, 16));
    }
    bf16* dst = y + (size_t)f * kTok * Ch + c0;
    f32x2 s2 = pk2(0.f, 0.f), q2 = pk2(0.f, 0.f);
#pragma unroll
    for (int r = 3; r >= 0; --r)
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        f32x2 acc = bias;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int lr = r + ky - 1;
          if (lr < 0) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int ix = x + kx - 1;
            if (ix >= 0 && ix <= 7) acc = fma2(a[lr][ix], w[ky * 3 + kx], acc);
          }
        }
        s2 = add2(s2, acc);
        q2 = fma2(acc, acc, q2);
        *reinterpret_cast<uint32_t*>(dst + (size_t)(pbase + r * pstep + x) * Ch) = f32x2_to_bf16x2(acc);
      }
    float s_lo, s_hi, q_lo, q_hi;
    upk2(s2, s_lo, s_hi);
    upk2(q2, q_lo, q_hi);
    const float s = warp_sum(s_lo + s_hi), q = warp_sum(q_lo + q_hi);
    if (lane == 0) { red[it][grp][gw] = s; red[it][grp][4 + gw] = q; }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(kFfnChunk) : "memory");
    if (tc == 0) {
      float* p = partial2 + ((size_t)f * gridDim.x + blockIdx.x) * 2;
      p[0] = (red[it][grp][0] + red[it][grp][1]) + (red[it][grp][2] + red[it][grp][3]);
      p[1] = (red[it][grp][4] + red[it][grp][5]) + (red[it][grp][6] + red[it][grp][7]);
    }
  }
}

static int g_opt_ffn_scalar = 0;                                  // npvp_set_option("ffn_scalar", 1): the scalar kernels (A/B runs)
void npvp_set_ffn_scalar(int v) { g_opt_ffn_scalar = v; }

static int launch_ffn_dwconv2(const void* h_bf16, const float* stats1, const float* n1w, const float* n1b, const float* dw_w,
                              const float* dw_b, void* y_bf16, float* partial2, int64_t frames, cudaStream_t st) {
  constexpr int smem = kTok * kDw2Pairs * (int)sizeof(float4);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(ffn_dwconv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) { npvp_set_error("ffn_dwconv2: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(err)); return NPVP_ERR_CUDA; }
    attr_set = true;
  }
  const int64_t chunks = 2048 / kFfnChunk;
  int64_t gy = 148 / chunks;                                     // one resident block per SM, groups stride over the frames
  const int64_t gy_max = (frames + kDw2Groups - 1) / kDw2Groups;
  gy = gy < 1 ? 1 : (gy > gy_max ? gy_max : gy);
  dim3 grid((unsigned)chunks, (unsigned)gy);
  ffn_dwconv2_kernel<<<grid, kFfnChunk * kDw2Groups, smem, st>>>((const bf16*)h_bf16, stats1, n1w, n1b, dw_w, dw_b, (bf16*)y_bf16, partial2,
                                                                 (int)frames);
  NPVP_LAUNCH_CHECK("ffn_dwconv2_kernel");
  return NPVP_OK;
}

template <int CH>
static int launch_ffn_dwconv(const void* h_bf16, const float* stats1, const float* n1w, const float* n1b, const float* dw_w,
                             const float* dw_b, void* y_bf16, float* partial2, int64_t frames, int64_t Ch, cudaStream_t st) {
  constexpr int smem = kTok * kFfnChunk * (int)sizeof(float2);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(ffn_dwconv_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) { npvp_set_error("ffn_dwconv: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(err)); return NPVP_ERR_CUDA; }
    attr_set = true;
  }
  // one resident 640-thread block per SM: a single wave of blocks, each group striding over the frames
  const int64_t chunks = Ch / kFfnChunk;
  int64_t gy = 148 / chunks;
  const int64_t gy_max = (frames + kFfnGroups - 1) / kFfnGroups;
  gy = gy < 1 ? 1 : (gy > gy_max ? gy_max : gy);
  dim3 grid((unsigned)chunks, (unsigned)gy);
  ffn_dwconv_kernel<CH><<<grid, kFfnChunk * kFfnGroups, smem, st>>>((const bf16*)h_bf16, stats1, n1w, n1b, dw_w, dw_b, (bf16*)y_bf16, partial2,
                                                                    (int)Ch, (int)frames);
  NPVP_LAUNCH_CHECK("ffn_dwconv_kernel");
  return NPVP_OK;
}

extern "C" int npvp_ffn_dwconv(const void* h_bf16, const float* stats1, const float* n1w, const float* n1b, const float* dw_w,
                               const float* dw_b, void* y_bf16, float* partial2, int64_t frames, int64_t Ch, void* stream) {
  NPVP_REQUIRE(h_bf16 && stats1 && n1w && n1b && dw_w && dw_b && y_bf16 && partial2, "npvp_ffn_dwconv: null pointer");
  NPVP_REQUIRE(frames > 0 && frames < (1ll << 31) && Ch % kFfnChunk == 0, "npvp_ffn_dwconv: frames > 0, Ch multiple of %d", kFfnChunk);
  cudaStream_t st = (cudaStream_t)stream;
  if (Ch == 2048 && !g_opt_ffn_scalar) return launch_ffn_dwconv2(h_bf16, stats1, n1w, n1b, dw_w, dw_b, y_bf16, partial2, frames, st);   // NPVP's only width
  if (Ch == 2048) return launch_ffn_dwconv<2048>(h_bf16, stats1, n1w, n1b, dw_w, dw_b, y_bf16, partial2, frames, Ch, st);   // NPVP's only width
  return launch_ffn_dwconv<0>(h_bf16, stats1, n1w, n1b, dw_w, dw_b, y_bf16, partial2, frames, Ch, st);
}

// out = GELU(LN2(y)).  Same parameter-traffic argument as above: a thread owns 8 channels of one pixel, keeps their 16
// LayerNorm parameters in registers and strides over frames blockIdx.y, blockIdx.y + gridDim.y, ...  The fp64 reduction of
// the per-chunk partial sums (one frame per thread) runs once per block and is kept in shared memory.
constexpr int kNorm2MaxFrames = 256;                             // frames per block (host sizes gridDim.y accordingly)
template <bool PACKED>
__global__ void __launch_bounds__(256, 4)
ffn_norm2_kernel(const bf16* __restrict__ y, const float* __restrict__ partial2, int nchunk, const float* __restrict__ n2w,
                 const float* __restrict__ n2b, bf16* __restrict__ out, int Ch, int frames) {
  __shared__ float2 s_stats[kNorm2MaxFrames];
  const int vec_per_px = Ch / 8, vblocks = (vec_per_px + 255) / 256;
  const int p = blockIdx.x / vblocks, cv = (blockIdx.x % vblocks) * 256 + threadIdx.x;
  {
    const int f = blockIdx.y + threadIdx.x * gridDim.y;
    if (f < frames) {
      double s = 0.0, q = 0.0;
      for (int k = 0; k < nchunk; ++k) {
        s += (double)__ldg(partial2 + ((size_t)f * nchunk + k) * 2);
        q += (double)__ldg(partial2 + ((size_t)f * nchunk + k) * 2 + 1);
      }
      const double n = (double)kTok * (double)Ch;
      const double mean_d = s / n;
      s_stats[threadIdx.x] = make_float2((float)mean_d, (float)(1.0 / sqrt(fmax(q / n - mean_d * mean_d, 0.0) + (double)kEps)));
    }
  }
  __syncthreads();
  if (cv >= vec_per_px) return;
  const size_t aoff = (size_t)p * Ch + (size_t)cv * 8;
  const float4 w0 = __ldg(reinterpret_cast<const float4*>(n2w + aoff)), w1 = __ldg(reinterpret_cast<const float4*>(n2w + aoff + 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(n2b + aoff)), b1 = __ldg(reinterpret_cast<const float4*>(n2b + aoff + 4));
  // The loop body is load -> ~80 dependent instructions -> store; left to the compiler a thread has ONE 16-byte load in
  // flight (r01 ncu: long-scoreboard stalls 13 per issued instruction, 2.2 TB/s read), so the loads run kN2Ahead frames
  // ahead of the arithmetic through a small register ring.
  constexpr int kN2Ahead = 3;
  const size_t base = (size_t)p * Ch + (size_t)cv * 8, fstride = (size_t)kTok * Ch;
  uint4 ring[kN2Ahead];
#pragma unroll
  for (int j = 0; j < kN2Ahead; ++j) {
    const int fj = blockIdx.y + j * gridDim.y;
    if (fj < frames) ring[j] = __ldg(reinterpret_cast<const uint4*>(y + (size_t)fj * fstride + base));
  }
  int k = 0;
  for (int f0 = blockIdx.y; f0 < frames; f0 += kN2Ahead * gridDim.y)
#pragma unroll
  for (int j = 0; j < kN2Ahead; ++j, ++k) {
    const int f = f0 + j * gridDim.y;
    if (f >= frames) break;
    const float2 st = s_stats[k];
    const float mean = st.x, rstd = st.y;
    const size_t off = (size_t)f * fstride + base;
    const uint4 u = ring[j];
    const int fn = f + kN2Ahead * gridDim.y;
    if (fn < frames) ring[j] = __ldg(reinterpret_cast<const uint4*>(y + (size_t)fn * fstride + base));
    uint4 o;
    if (PACKED) {                                                // FFMA2: (y * rstd - mean * rstd) * w + b as two packed FMAs per pair
      const float nmr = -mean * rstd;
      const f32x2 rs = pk2(rstd, rstd), nm = pk2(nmr, nmr);
      o.x = f32x2_to_bf16x2(gelu_erf2(fma2(fma2(bf16x2_to_f32x2(u.x), rs, nm), pk2(w0.x, w0.y), pk2(b0.x, b0.y))));
      o.y = f32x2_to_bf16x2(gelu_erf2(fma2(fma2(bf16x2_to_f32x2(u.y), rs, nm), pk2(w0.z, w0.w), pk2(b0.z, b0.w))));
      o.z = f32x2_to_bf16x2(gelu_erf2(fma2(fma2(bf16x2_to_f32x2(u.z), rs, nm), pk2(w1.x, w1.y), pk2(b1.x, b1.y))));
      o.w = f32x2_to_bf16x2(gelu_erf2(fma2(fma2(bf16x2_to_f32x2(u.w), rs, nm), pk2(w1.z, w1.w), pk2(b1.z, b1.w))));
    } else {
      const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
      o.x = pack_bf16x2(gelu_erf((p0.x - mean) * rstd * w0.x + b0.x), gelu_erf((p0.y - mean) * rstd * w0.y + b0.y));
      o.y = pack_bf16x2(gelu_erf((p1.x - mean) * rstd * w0.z + b0.z), gelu_erf((p1.y - mean) * rstd * w0.w + b0.w));
      o.z = pack_bf16x2(gelu_erf((p2.x - mean) * rstd * w1.x + b1.x), gelu_erf((p2.y - mean) * rstd * w1.y + b1.y));
      o.w = pack_bf16x2(gelu_erf((p3.x - mean) * rstd * w1.z + b1.z), gelu_erf((p3.y - mean) * rstd * w1.w + b1.w));
    }
    *reinterpret_cast<uint4*>(out + off) = o;
  }
}

extern "C" int npvp_ffn_norm2(const void* y_bf16, const float* partial2, const float* n2w, const float* n2b, void* out_bf16,
                              int64_t frames, int64_t Ch, void* stream) {
  NPVP_REQUIRE(y_bf16 && partial2 && n2w && n2b && out_bf16, "npvp_ffn_norm2: null pointer");
  NPVP_REQUIRE(frames > 0 && frames < (1ll << 31) && Ch % kFfnChunk == 0, "npvp_ffn_norm2: frames > 0, Ch multiple of %d", kFfnChunk);
  const int64_t gx = kTok * ((Ch / 8 + 255) / 256);
  int64_t gy = (148 * 4) / gx;                                   // one wave of 4 resident blocks per SM (58 registers) ...
  gy = gy < 1 ? 1 : gy;
  const int64_t gy_min = (frames + kNorm2MaxFrames - 1) / kNorm2MaxFrames;   // ... unless a block would exceed its stats table
  gy = gy < gy_min ? gy_min : gy;
  gy = gy > frames ? frames : gy;
  NPVP_REQUIRE(gy <= 65535, "npvp_ffn_norm2: too many frames per launch (%lld)", (long long)frames);
  dim3 grid((unsigned)gx, (unsigned)gy);
  if (g_opt_ffn_scalar)
    ffn_norm2_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)y_bf16, partial2, (int)(Ch / kFfnChunk), n2w, n2b,
                                                                   (bf16*)out_bf16, (int)Ch, (int)frames);
  else
    ffn_norm2_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)y_bf16, partial2, (int)(Ch / kFfnChunk), n2w, n2b,
                                                                  (bf16*)out_bf16, (int)Ch, (int)frames);
  NPVP_LAUNCH_CHECK("ffn_norm2_kernel");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// Fourier features
// ---------------------------------------------------------------------------------------------
__global__ void fourier_features_kernel(const float* __restrict__ coor, const float* __restrict__ B, float* __restrict__ out,
                                        int64_t rows, int half) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * half) return;
  const int64_t r = i / half;
  const int j = (int)(i % half);
  const float two_pi = 2.0f * 3.14159265358979323846f;   // matches 2. * float(math.pi) rounded to fp32 operands
  const float x0 = two_pi * coor[r * 3], x1 = two_pi * coor[r * 3 + 1], x2 = two_pi * coor[r * 3 + 2];
  float p = x0 * B[j * 3];
  p = fmaf(x1, B[j * 3 + 1], p);
  p = fmaf(x2, B[j * 3 + 2], p);
  out[r * 2 * half + j] = cosf(p);
  out[r * 2 * half + half + j] = sinf(p);
}

extern "C" int npvp_fourier_features(const float* coor, const float* B, float* out, int64_t rows, int half, void* stream) {
  NPVP_REQUIRE(coor && B && out && rows > 0 && half > 0, "npvp_fourier_features: bad arguments");
  fourier_features_kernel<<<(unsigned)ceil_div64(rows * half, 256), 256, 0, (cudaStream_t)stream>>>(coor, B, out, rows, half);
  NPVP_LAUNCH_CHECK("fourier_features_kernel");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// event encoder pieces
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dwconv3x3_tokens_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ shift,
                        bf16* __restrict__ out, int C, int relu) {
  const int f = blockIdx.y, c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  float a[64], wk[9];
  const float* src = x + (size_t)f * kTok * C + c;
#pragma unroll
  for (int p = 0; p < 64; ++p) a[p] = __ldg(src + (size_t)p * C);
#pragma unroll
  for (int k = 0; k < 9; ++k) wk[k] = __ldg(w + (size_t)k * C + c);
  const float sh = __ldg(shift + c);
  bf16* dst = out + (size_t)f * kTok * C + c;
#pragma unroll
  for (int yy = 0; yy < 8; ++yy)
#pragma unroll
    for (int xx = 0; xx < 8; ++xx) {
      float o = dw3x3_at(a, wk, sh, yy, xx);
      if (relu) o = fmaxf(o, 0.f);
      dst[(size_t)(yy * 8 + xx) * C] = __float2bfloat16(o);
    }
}

extern "C" int npvp_dwconv3x3_tokens(const float* x, const float* w, const float* shift, void* out_bf16, int64_t frames,
                                     int64_t C, int relu, void* stream) {
  NPVP_REQUIRE(x && w && shift && out_bf16 && frames > 0 && frames <= 65535 && C > 0, "npvp_dwconv3x3_tokens: bad arguments");
  dim3 grid((unsigned)ceil_div64(C, 256), (unsigned)frames);
  dwconv3x3_tokens_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, shift, (bf16*)out_bf16, (int)C, relu);
  NPVP_LAUNCH_CHECK("dwconv3x3_tokens_kernel");
  return NPVP_OK;
}

__global__ void latent_reparam_kernel(const float* __restrict__ mulv, int64_t ld, const float* __restrict__ eps,
                                      float* __restrict__ z, int64_t n_clips, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_clips * kTok * C) return;
  const int c = (int)(i % C);
  const int64_t tok = i / C;           // n*64 + p
  const int64_t n = tok / kTok;
  const int p = (int)(tok % kTok);
  const float mu = mulv[tok * ld + c];
  float v = mu;
  if (eps) {
    const float lv = mulv[tok * ld + C + c];
    v = fmaf(expf(0.5f * lv), eps[(n * C + c) * kTok + p], mu);
  }
  z[i] = v;
}

extern "C" int npvp_latent_reparam(const float* mulv, int64_t ld, const float* eps_nchw, float* z, int64_t n_clips, int64_t C,
                                   void* stream) {
  NPVP_REQUIRE(mulv && z && n_clips > 0 && C > 0 && ld >= (eps_nchw ? 2 * C : C), "npvp_latent_reparam: bad arguments");
  const int64_t total = n_clips * kTok * C;
  latent_reparam_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(mulv, ld, eps_nchw, z, n_clips, (int)C);
  NPVP_LAUNCH_CHECK("latent_reparam_kernel");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// layout changes: per frame [C, HW] <-> [HW, C] through 32x32 smem tiles
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nchw_to_tokens_kernel(const float* __restrict__ x, float* __restrict__ out_f32, h16* __restrict__ out_bf16, int C, int HW, int fp16) {
  __shared__ float tile[32][33];
  const size_t base = (size_t)blockIdx.z * C * HW;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    tile[r][tx] = (c < C && p < HW) ? __ldg(x + base + (size_t)c * HW + p) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    if (p < HW && c < C) {
      const float v = tile[tx][r];
      if (out_f32) out_f32[base + (size_t)p * C + c] = v;
      if (out_bf16) out_bf16[base + (size_t)p * C + c] = float_to_h16(v, fp16);
    }
  }
}

extern "C" int npvp_nchw_to_tokens(const float* x, float* out_f32, void* out_bf16, int64_t frames, int64_t C, int64_t HW, int fp16, void* stream) {
  NPVP_REQUIRE(x && (out_f32 || out_bf16) && frames > 0 && frames <= 65535 && C > 0 && HW > 0, "npvp_nchw_to_tokens: bad arguments");
  dim3 grid((unsigned)ceil_div64(HW, 32), (unsigned)ceil_div64(C, 32), (unsigned)frames);
  nchw_to_tokens_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out_f32, (h16*)out_bf16, (int)C, (int)HW, fp16);
  NPVP_LAUNCH_CHECK("nchw_to_tokens_kernel");
  return NPVP_OK;
}

__global__ void __launch_bounds__(256)
tokens_to_nchw_kernel(const float* __restrict__ x_f32, const h16* __restrict__ x_bf16, float* __restrict__ out, int C, int HW, int relu, int fp16) {
  __shared__ float tile[32][33];
  const size_t base = (size_t)blockIdx.z * C * HW;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    float v = 0.f;
    if (p < HW && c < C) v = x_f32 ? __ldg(x_f32 + base + (size_t)p * C + c) : h16_to_float(x_bf16[base + (size_t)p * C + c], fp16);
    tile[r][tx] = relu ? fmaxf(v, 0.f) : v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    if (c < C && p < HW) out[base + (size_t)c * HW + p] = tile[tx][r];
  }
}

extern "C" int npvp_tokens_to_nchw(const float* x_f32, const void* x_bf16, float* out, int64_t frames, int64_t C, int64_t HW,
                                   int relu, int fp16, void* stream) {
  NPVP_REQUIRE((x_f32 || x_bf16) && out && frames > 0 && frames <= 65535 && C > 0 && HW > 0, "npvp_tokens_to_nchw: bad arguments");
  dim3 grid((unsigned)ceil_div64(HW, 32), (unsigned)ceil_div64(C, 32), (unsigned)frames);
  tokens_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_f32, (const h16*)x_bf16, out, (int)C, (int)HW, relu, fp16);
  NPVP_LAUNCH_CHECK("tokens_to_nchw_kernel");
  return NPVP_OK;
}
