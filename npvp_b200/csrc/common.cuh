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

// nn.GELU() default (erf form): x * Phi(x).  erf via Abramowitz-Stegun 7.1.26 (one rcp + one ex2 + 5 FMA): measured
// max |gelu - exact| = 4.1e-7 over [-8, 8], i.e. below fp32 round-off of the surrounding LayerNorms, at ~22 instructions
// instead of ~38 for erff().  The conv-FFN evaluates 2 x 131072 GELUs per frame, which made erff the single largest
// instruction consumer of the memory-bound kernels (r01 profile).
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float h = 0.5f * p * t * ex2_approx(-z * z * 1.4426950408889634f);   // 0.5 * erfc(|x| / sqrt 2)
  return x * (x >= 0.f ? 1.0f - h : h);
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
  if (fp16) {
    __half2 v = __floats2half2_rn(fminf(fmaxf(lo, -65504.f), 65504.f), fminf(fmaxf(hi, -65504.f), 65504.f));
    return *reinterpret_cast<uint32_t*>(&v);
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
  int64_t ld_out, ld_res;
};

__device__ __forceinline__ float epi_value(const EpiParams& e, float acc, int64_t m, int64_t n) {
  float v = acc;
  if (e.bias) v += __ldg(e.bias + n);
  v = apply_act(v, e.act);
  v *= e.alpha;
  if (e.res1) v += e.res1_bf16 ? h16_to_float(((const h16*)e.res1)[m * e.ld_res + n], e.fp16) : ((const float*)e.res1)[m * e.ld_res + n];
  if (e.res2) v += e.res2_bf16 ? h16_to_float(((const h16*)e.res2)[m * e.ld_res + n], e.fp16) : ((const float*)e.res2)[m * e.ld_res + n];
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
  e.alpha = ep->alpha;
  e.act = ep->act;
  e.res1_bf16 = ep->res1_bf16;
  e.res2_bf16 = ep->res2_bf16;
  e.post_relu = ep->post_relu;
  e.ld_out = ep->ld_out;
  e.ld_res = ep->ld_res;
  return e;
}
