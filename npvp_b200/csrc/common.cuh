// Common device/host helpers for the npvp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/npvp_b200.h"

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// error plumbing: every extern "C" entry returns 0 on success, <0 on failure; message via npvp_last_error()
// ---------------------------------------------------------------------------------------------
void npvp_set_error(const char* fmt, ...);
int npvp_check_launch(const char* what);
extern int g_launch_count;   // number of kernel launches issued through this library (for bench accounting)

#define NPVP_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) {                                                \
      npvp_set_error(__VA_ARGS__);                                \
      return NPVP_ERR_INVALID;                                    \
    }                                                             \
  } while (0)

#define NPVP_LAUNCH_CHECK(name)                                   \
  do {                                                            \
    ++g_launch_count;                                             \
    int _rc = npvp_check_launch(name);                            \
    if (_rc) return _rc;                                          \
  } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// nn.GELU() default (erf form): x * Phi(x) = max(x, 0) - |x| * Q(|x|), Q(u) = 0.5 erfc(u / sqrt 2) the normal upper tail.
// log2 Q(u) is smooth on [0, 5.5], so Q = ex2(P6(u)) with a degree-6 polynomial fitted (Lawson-weighted minimax on the
// GELU error u Q ln2 dP) to max |gelu - exact| = 2.8e-7 in fp32 over [-9, 9]; beyond 5.5 the tail term is < 1.1e-7 and u is
// clamped.  10 instructions, one MUFU -- the previous Abramowitz-Stegun 7.1.26 form took ~22 with two MUFU and erff() ~38.
// The conv-FFN evaluates 2 x 131072 GELUs per frame, which made GELU the largest instruction consumer of the
// memory-bound kernels (r01 profile).  max.NaN keeps NaN inputs visible.
__device__ __forceinline__ float gelu_erf(float x) {
  const float u = fminf(fabsf(x), 5.5f);
  float p = fmaf(3.309290792e-05f, u, -7.692205073e-04f);
  p = fmaf(p, u, 8.080719144e-03f);
  p = fmaf(p, u, -5.341210813e-02f);
  p = fmaf(p, u, -4.587709705e-01f);
  p = fmaf(p, u, -1.151201703e+00f);
  p = fmaf(p, u, -9.999930609e-01f);
  float r;
  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(x));
  return fmaf(-u, ex2_approx(p), r);
}

// the same GELU on 8 values with the 8 Horner chains interleaved step by step (for kernels with few warps per scheduler)
__device__ __forceinline__ void gelu_erf_x8(float (&x)[8]) {
  float u[8], p[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { u[j] = fminf(fabsf(x[j]), 5.5f); p[j] = fmaf(3.309290792e-05f, u[j], -7.692205073e-04f); }
#pragma unroll
  for (int j = 0; j < 8; ++j) p[j] = fmaf(p[j], u[j], 8.080719144e-03f);
#pragma unroll
  for (int j = 0; j < 8; ++j) p[j] = fmaf(p[j], u[j], -5.341210813e-02f);
#pragma unroll
  for (int j = 0; j < 8; ++j) p[j] = fmaf(p[j], u[j], -4.587709705e-01f);
#pragma unroll
  for (int j = 0; j < 8; ++j) p[j] = fmaf(p[j], u[j], -1.151201703e+00f);
#pragma unroll
  for (int j = 0; j < 8; ++j) p[j] = fmaf(p[j], u[j], -9.999930609e-01f);
#pragma unroll
  for (int j = 0; j < 8; ++j) p[j] = ex2_approx(p[j]);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float r;
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(x[j]));
    x[j] = fmaf(-u[j], p[j], r);
  }
}

// Packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2: one instruction, two IEEE-rn fp32 results).  Measured on B200
// (tools/ubench/ffma2.cu): the FMA pipe retires the same 128 FMA / clk / SM either way, so packing does not add arithmetic
// throughput - it halves the ISSUE slots the arithmetic takes, which is what bounds the elementwise-heavy kernels here
// (LayerNorm affine + GELU + depthwise taps: ~19 of ~27 issued instructions per element are FMA-pipe).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 bf16x2_to_f32x2(uint32_t u) { return pk2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
__device__ __forceinline__ uint32_t f32x2_to_bf16x2(f32x2 v) {
  float lo, hi;
  upk2(v, lo, hi);
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// gelu_erf on a packed pair, bit-identical to the scalar form: the Horner chain runs in nu = -min(|x|, 5.5) with the odd
// coefficients negated (round-to-nearest is sign symmetric), so the final step is fma(nu, Q, max(x, 0)) with no negation.
// Per pair: 2 FMNMX + 6 FFMA2 + 2 MUFU + 2 FMNMX + 1 FFMA2 = 13 issue slots (20 scalar).
__device__ __forceinline__ f32x2 gelu_erf2(f32x2 x) {
  float x0, x1;
  upk2(x, x0, x1);
  const f32x2 nu = pk2(fmaxf(-fabsf(x0), -5.5f), fmaxf(-fabsf(x1), -5.5f));
  f32x2 p = fma2(pk2(3.309290792e-05f, 3.309290792e-05f), nu, pk2(7.692205073e-04f, 7.692205073e-04f));
  p = fma2(p, nu, pk2(8.080719144e-03f, 8.080719144e-03f));
  p = fma2(p, nu, pk2(5.341210813e-02f, 5.341210813e-02f));
  p = fma2(p, nu, pk2(-4.587709705e-01f, -4.587709705e-01f));
  p = fma2(p, nu, pk2(1.151201703e+00f, 1.151201703e+00f));
  p = fma2(p, nu, pk2(-9.999930609e-01f, -9.999930609e-01f));
  float p0, p1, r0, r1;
  upk2(p, p0, p1);
  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r0) : "f"(x0));
  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r1) : "f"(x1));
  return fma2(nu, pk2(ex2_approx(p0), ex2_approx(p1)), pk2(r0, r1));
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == NPVP_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == NPVP_ACT_GELU) return gelu_erf(v);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of two values (sum, sumsq).  `red` must hold 2*32 floats.  All threads get the result.
__device__ __forceinline__ void block_sum2(float& a, float& b, float* red) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();                       // protect `red` from a previous use
  if (lane == 0) { red[warp] = a; red[32 + warp] = b; }
  __syncthreads();
  float x = (lane < nw) ? red[lane] : 0.0f;
  float y = (lane < nw) ? red[32 + lane] : 0.0f;
  a = warp_sum(x);
  b = warp_sum(y);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// 16-bit storage type chosen at run time: fp16 = 0 -> bfloat16, fp16 = 1 -> IEEE half (saturating stores).
// The autoencoder runs in half (its errors go straight to pixels), the predictor in bfloat16 (range safety).
typedef uint16_t h16;
__device__ __forceinline__ uint32_t pack_h16x2(float lo, float hi, int fp16) {
  if (fp16) {                                  // one F2FP.SATFINITE: round to nearest, |x| > 65504 -> +-65504 (was 4 FMNMX + F2FP)
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  return pack_bf16x2(lo, hi);
}
__device__ __forceinline__ float2 unpack_h16x2(uint32_t u, int fp16) {
  if (fp16) return __half22float2(*reinterpret_cast<__half2*>(&u));
  return unpack_bf16x2(u);
}
__device__ __forceinline__ float h16_to_float(h16 v, int fp16) {
  if (fp16) return __half2float(*reinterpret_cast<__half*>(&v));
  return __uint_as_float((uint32_t)v << 16);
}
__device__ __forceinline__ h16 float_to_h16(float v, int fp16) {
  if (fp16) { __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); return *reinterpret_cast<h16*>(&h); }
  bf16 b = __float2bfloat16(v);
  return *reinterpret_cast<h16*>(&b);
}

// ---------------------------------------------------------------------------------------------
// shared GEMM epilogue: v = acc (+bias[n]) -> act -> *alpha -> +res1 -> +res2 -> relu? ; store f32 / bf16
// ---------------------------------------------------------------------------------------------
struct EpiParams {
  const float* bias;
  const void* res1;
  const void* res2;
  float* out_f32;
  h16* out_bf16;      // 16-bit output (bf16 or half, see fp16)
  float alpha;
  int act, res1_bf16, res2_bf16, post_relu, fp16;
  int out_fp16;        // 16-bit type of out_bf16 and of 16-bit residuals (npvp_epilogue_t.out16: defaults to the operands' type)
  int64_t ld_out, ld_res;
  float* frame_stats;  // see npvp_epilogue_t
};

__device__ __forceinline__ float epi_value(const EpiParams& e, float acc, int64_t m, int64_t n) {
  float v = acc;
  if (e.bias) v += __ldg(e.bias + n);
  v = apply_act(v, e.act);
  v *= e.alpha;
  if (e.res1) v += e.res1_bf16 ? h16_to_float(((const h16*)e.res1)[m * e.ld_res + n], e.out_fp16) : ((const float*)e.res1)[m * e.ld_res + n];
  if (e.res2) v += e.res2_bf16 ? h16_to_float(((const h16*)e.res2)[m * e.ld_res + n], e.out_fp16) : ((const float*)e.res2)[m * e.ld_res + n];
  if (e.post_relu) v = fmaxf(v, 0.0f);
  return v;
}

static inline EpiParams make_epi(const npvp_epilogue_t* ep) {
  EpiParams e;
  e.bias = (const float*)ep->bias;
  e.res1 = ep->res1;
  e.res2 = ep->res2;
  e.out_f32 = (float*)ep->out_f32;
  e.out_bf16 = (h16*)ep->out_bf16;
  e.fp16 = ep->fp16;
  e.out_fp16 = ep->out16 == 0 ? ep->fp16 : (ep->out16 == 1 ? 1 : 0);
  e.alpha = ep->alpha;
  e.act = ep->act;
  e.res1_bf16 = ep->res1_bf16;
  e.res2_bf16 = ep->res2_bf16;
  e.post_relu = ep->post_relu;
  e.ld_out = ep->ld_out;
  e.ld_res = ep->ld_res;
  e.frame_stats = ep->frame_stats;
  return e;
}
