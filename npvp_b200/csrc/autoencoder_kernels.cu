// Autoencoder kernels that are not plain GEMMs: the 7x7 reflect-padded stem and head convolutions, the
// patch gather that turns 3x3 / strided / transposed convolutions into tensor-core GEMMs, 2x2 max-pooling of
// the non-local key/value projections and the non-local attention core (unscaled softmax).
// Activations are bf16 channels-last (NHWC).  A transposed conv is computed as a GEMM over the 2x2 input
// neighbourhood whose N dimension enumerates the 4 output phases, so its output is "phase-major"
// [frames, H/2, W/2, (py,px), C]; consumers take a `phase_major` flag instead of paying for a pixel shuffle.
#include "common.cuh"

__device__ __forceinline__ int reflect_idx(int i, int n) {   // ReflectionPad2d: mirror without repeating the edge
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

__device__ __forceinline__ size_t pixel_offset(int64_t f, int y, int x, int H, int W, int phase_major) {
  if (phase_major) return (((size_t)f * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * 4 + ((y & 1) << 1) + (x & 1);
  return ((size_t)f * H + y) * W + x;
}

// ---------------------------------------------------------------------------------------------
// stem: fp32 NCHW image -> bf16 NHWC features.  block = 16x16 pixels, thread = 1 pixel x 16 output channels
// ---------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256)
conv7x7_stem_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ shift,
                    h16* __restrict__ out, int Cout, int H, int W, int fp16) {
  __shared__ float tile[CIN][22][23];
  __shared__ __align__(16) float ws[49 * CIN][16];
  const int f = blockIdx.z / (Cout / 16), cg = blockIdx.z % (Cout / 16);
  const int tiles_x = (W + 15) / 16;
  const int ty0 = (blockIdx.x / tiles_x) * 16, tx0 = (blockIdx.x % tiles_x) * 16;
  for (int i = threadIdx.x; i < CIN * 22 * 22; i += 256) {
    const int c = i / (22 * 22), r = (i / 22) % 22, col = i % 22;
    const int yy = reflect_idx(ty0 + r - 3, H), xx = reflect_idx(tx0 + col - 3, W);
    tile[c][r][col] = __ldg(x + (((size_t)f * CIN + c) * H + yy) * W + xx);
  }
  for (int i = threadIdx.x; i < 49 * CIN * 16; i += 256) ws[i / 16][i % 16] = __ldg(w + (size_t)(i / 16) * Cout + cg * 16 + (i % 16));
  __syncthreads();
  const int ly = threadIdx.x / 16, lx = threadIdx.x % 16;
  const int oy = ty0 + ly, ox = tx0 + lx;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (int ky = 0; ky < 7; ++ky)
    for (int kx = 0; kx < 7; ++kx)
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const float v = tile[c][ly + ky][lx + kx];
        const float4* wr = reinterpret_cast<const float4*>(&ws[(ky * 7 + kx) * CIN + c][0]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 q = wr[j];
          acc[4 * j] = fmaf(v, q.x, acc[4 * j]); acc[4 * j + 1] = fmaf(v, q.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(v, q.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(v, q.w, acc[4 * j + 3]);
        }
      }
  if (oy < H && ox < W) {
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      pk[j] = pack_h16x2(fmaxf(acc[2 * j] + __ldg(shift + cg * 16 + 2 * j), 0.f), fmaxf(acc[2 * j + 1] + __ldg(shift + cg * 16 + 2 * j + 1), 0.f), fp16);
    uint4* dst = reinterpret_cast<uint4*>(out + (((size_t)f * H + oy) * W + ox) * Cout + cg * 16);
    dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

extern "C" int npvp_conv7x7_stem(const float* x, const float* w, const float* shift, void* out_bf16, int64_t frames, int Cin,
                                 int Cout, int H, int W, int fp16, void* stream) {
  NPVP_REQUIRE(x && w && shift && out_bf16 && frames > 0, "npvp_conv7x7_stem: bad arguments");
  NPVP_REQUIRE(Cout % 16 == 0 && H >= 4 && W >= 4, "npvp_conv7x7_stem: Cout must be a multiple of 16, H/W >= 4");
  NPVP_REQUIRE(frames * (Cout / 16) <= 65535, "npvp_conv7x7_stem: too many frames per launch (%lld)", (long long)frames);
  dim3 grid((unsigned)(((H + 15) / 16) * ((W + 15) / 16)), 1, (unsigned)(frames * (Cout / 16)));
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 1) conv7x7_stem_kernel<1><<<grid, 256, 0, st>>>(x, w, shift, (h16*)out_bf16, Cout, H, W, fp16);
  else if (Cin == 3) conv7x7_stem_kernel<3><<<grid, 256, 0, st>>>(x, w, shift, (h16*)out_bf16, Cout, H, W, fp16);
  else NPVP_REQUIRE(false, "npvp_conv7x7_stem: Cin must be 1 or 3 (got %d)", Cin);
  NPVP_LAUNCH_CHECK("conv7x7_stem_kernel");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// head: bf16 NHWC (or phase-major) features -> fp32 NCHW image, + bias + tanh/sigmoid.
// block = 16 rows x 64 cols of output; thread = 4 adjacent pixels x COUT channels; input staged 4 channels at a time.
// ---------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(256)
conv7x7_head_kernel(const h16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                    int Cin, int H, int W, int phase_major, int act, int fp16) {
  constexpr int CCH = 4;
  __shared__ __align__(16) float tile[CCH][22][72];          // 70 used columns, padded to 72
  __shared__ float ws[CCH][49][COUT];
  const int f = blockIdx.z;
  const int tiles_x = (W + 63) / 64;
  const int ty0 = (blockIdx.x / tiles_x) * 16, tx0 = (blockIdx.x % tiles_x) * 64;
  const int ly = threadIdx.x / 16, lx4 = (threadIdx.x % 16) * 4;
  float acc[4][COUT];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[p][co] = 0.f;
  for (int c0 = 0; c0 < Cin; c0 += CCH) {
    __syncthreads();
    for (int i = threadIdx.x; i < 22 * 70; i += 256) {
      const int r = i / 70, col = i % 70;
      const int yy = reflect_idx(ty0 + r - 3, H), xx = reflect_idx(tx0 + col - 3, W);
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(x + pixel_offset(f, yy, xx, H, W, phase_major) * Cin + c0));
      const float2 a = unpack_h16x2(u.x, fp16), b = unpack_h16x2(u.y, fp16);
      tile[0][r][col] = a.x; tile[1][r][col] = a.y; tile[2][r][col] = b.x; tile[3][r][col] = b.y;
    }
    for (int i = threadIdx.x; i < CCH * 49 * COUT; i += 256) {
      const int c = i / (49 * COUT), t = (i / COUT) % 49, co = i % COUT;
      ws[c][t][co] = __ldg(w + ((size_t)t * Cin + c0 + c) * COUT + co);
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CCH; ++c)
#pragma unroll 1
      for (int ky = 0; ky < 7; ++ky) {
        float in[10];
        const float* row = &tile[c][ly + ky][lx4];
        const float4 q0 = *reinterpret_cast<const float4*>(row), q1 = *reinterpret_cast<const float4*>(row + 4);
        const float2 q2 = *reinterpret_cast<const float2*>(row + 8);
        in[0] = q0.x; in[1] = q0.y; in[2] = q0.z; in[3] = q0.w; in[4] = q1.x; in[5] = q1.y; in[6] = q1.z; in[7] = q1.w;
        in[8] = q2.x; in[9] = q2.y;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx)
#pragma unroll
          for (int co = 0; co < COUT; ++co) {
            const float wv = ws[c][ky * 7 + kx][co];
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[p][co] = fmaf(in[p + kx], wv, acc[p][co]);
          }
      }
  }
  const int oy = ty0 + ly;
  if (oy < H) {
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      const float b = __ldg(bias + co);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int ox = tx0 + lx4 + p;
        if (ox < W) {
          float v = acc[p][co] + b;
          v = (act == NPVP_ACT_TANH) ? tanhf(v) : (act == NPVP_ACT_SIGMOID ? 1.0f / (1.0f + expf(-v)) : v);
          out[(((size_t)f * COUT + co) * H + oy) * W + ox] = v;
        }
      }
    }
  }
}

extern "C" int npvp_conv7x7_head(const void* x_bf16, const float* w, const float* bias, float* out, int64_t frames, int Cin,
                                 int Cout, int H, int W, int phase_major, int act, int fp16, void* stream) {
  NPVP_REQUIRE(x_bf16 && w && bias && out && frames > 0 && frames <= 65535, "npvp_conv7x7_head: bad arguments");
  NPVP_REQUIRE(Cin % 4 == 0 && H >= 4 && W >= 4 && (!phase_major || (H % 2 == 0 && W % 2 == 0)), "npvp_conv7x7_head: Cin %% 4, H/W >= 4, even H/W for phase-major input");
  dim3 grid((unsigned)(((H + 15) / 16) * ((W + 63) / 64)), 1, (unsigned)frames);
  cudaStream_t st = (cudaStream_t)stream;
  if (Cout == 1) conv7x7_head_kernel<1><<<grid, 256, 0, st>>>((const h16*)x_bf16, w, bias, out, Cin, H, W, phase_major, act, fp16);
  else if (Cout == 3) conv7x7_head_kernel<3><<<grid, 256, 0, st>>>((const h16*)x_bf16, w, bias, out, Cin, H, W, phase_major, act, fp16);
  else NPVP_REQUIRE(false, "npvp_conv7x7_head: Cout must be 1 or 3 (got %d)", Cout);
  NPVP_LAUNCH_CHECK("conv7x7_head_kernel");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// patch gather (im2col) in 16-byte vectors of 8 channels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col_nhwc_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int64_t total_vec, int H, int W, int C, int KH, int KW,
                   int stride, int pad, int pad_mode, int Ho, int Wo, int phase_major) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int cv = C >> 3;
  const int c8 = (int)(i % cv);
  int64_t rest = i / cv;
  const int tap = (int)(rest % (KH * KW));
  rest /= (KH * KW);
  const int ox = (int)(rest % Wo);
  rest /= Wo;
  const int oy = (int)(rest % Ho);
  const int64_t f = rest / Ho;
  int iy = oy * stride - pad + tap / KW, ix = ox * stride - pad + tap % KW;
  bool valid = true;
  if (pad_mode == NPVP_PAD_REFLECT) { iy = reflect_idx(iy, H); ix = reflect_idx(ix, W); }
  else if (pad_mode == NPVP_PAD_REPLICATE) { iy = min(max(iy, 0), H - 1); ix = min(max(ix, 0), W - 1); }
  else valid = (iy >= 0 && iy < H && ix >= 0 && ix < W);
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (valid) v = __ldg(reinterpret_cast<const uint4*>(x + pixel_offset(f, iy, ix, H, W, phase_major) * C) + c8);
  reinterpret_cast<uint4*>(out)[i] = v;
}

extern "C" int npvp_im2col_nhwc(const void* x_bf16, void* out_bf16, int64_t frames, int H, int W, int C, int KH, int KW, int stride,
                                int pad, int pad_mode, int Ho, int Wo, int phase_major, void* stream) {
  NPVP_REQUIRE(x_bf16 && out_bf16 && frames > 0, "npvp_im2col_nhwc: bad arguments");
  NPVP_REQUIRE(C % 8 == 0 && KH > 0 && KW > 0 && stride > 0 && Ho > 0 && Wo > 0, "npvp_im2col_nhwc: C must be a multiple of 8");
  NPVP_REQUIRE(!phase_major || (H % 2 == 0 && W % 2 == 0), "npvp_im2col_nhwc: phase-major input needs even H, W");
  NPVP_REQUIRE(pad_mode == NPVP_PAD_ZERO || (pad < H && pad < W), "npvp_im2col_nhwc: reflect/replicate pad must be smaller than the image");
  const int64_t total = frames * Ho * Wo * KH * KW * (C / 8);
  im2col_nhwc_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x_bf16, (bf16*)out_bf16, total, H, W, C, KH, KW,
                                                                                    stride, pad, pad_mode, Ho, Wo, phase_major);
  NPVP_LAUNCH_CHECK("im2col_nhwc_kernel");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// 2x2 max-pool of a column slice
// ---------------------------------------------------------------------------------------------
__global__ void maxpool2x2_cols_kernel(const h16* __restrict__ x, int64_t ldx, int col0, int Cn, h16* __restrict__ out,
                                       int64_t total, int H, int W, int fp16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % Cn);
  int64_t rest = i / Cn;
  const int Wo = W >> 1, Ho = H >> 1;
  const int ox = (int)(rest % Wo);
  rest /= Wo;
  const int oy = (int)(rest % Ho);
  const int64_t f = rest / Ho;
  const h16* p = x + (((size_t)f * H + 2 * oy) * W + 2 * ox) * ldx + col0 + c;
  const float a = h16_to_float(p[0], fp16), b = h16_to_float(p[ldx], fp16);
  const float d = h16_to_float(p[(size_t)W * ldx], fp16), e = h16_to_float(p[(size_t)W * ldx + ldx], fp16);
  out[i] = float_to_h16(fmaxf(fmaxf(a, b), fmaxf(d, e)), fp16);
}

extern "C" int npvp_maxpool2x2_cols(const void* x_bf16, int64_t ldx, int col0, int Cn, void* out_bf16, int64_t frames, int H, int W,
                                    int fp16, void* stream) {
  NPVP_REQUIRE(x_bf16 && out_bf16 && frames > 0 && Cn > 0 && col0 >= 0 && H % 2 == 0 && W % 2 == 0, "npvp_maxpool2x2_cols: bad arguments");
  const int64_t total = frames * (H / 2) * (W / 2) * Cn;
  maxpool2x2_cols_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>((const h16*)x_bf16, ldx, col0, Cn, (h16*)out_bf16, total, H, W, fp16);
  NPVP_LAUNCH_CHECK("maxpool2x2_cols_kernel");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// non-local attention core: thread = (query, 32-wide slice of dv); keys/values streamed through smem
// ---------------------------------------------------------------------------------------------
template <int DQ>
__global__ void __launch_bounds__(256)
nonlocal_attention_kernel(const h16* __restrict__ q, int64_t ldq, const h16* __restrict__ kv, h16* __restrict__ out, int HW, int HWk, int fp16) {
  constexpr int DV = 4 * DQ, S = DV / 32, QB = 256 / S, KT = 64, ROW = DQ + DV;
  __shared__ __align__(16) h16 skv[KT][ROW];
  const int blocks_per_frame = (HW + QB - 1) / QB;
  const int64_t f = blockIdx.x / blocks_per_frame;
  const int qi = (blockIdx.x % blocks_per_frame) * QB + (threadIdx.x % QB);
  const int slice = threadIdx.x / QB;
  const bool active = qi < HW;
  float qr[DQ];
  if (active) {
    const h16* qp = q + ((size_t)f * HW + qi) * ldq;
#pragma unroll
    for (int d = 0; d < DQ; d += 2) {
      const float2 t = unpack_h16x2(*reinterpret_cast<const uint32_t*>(qp + d), fp16);
      qr[d] = t.x * 1.4426950408889634f; qr[d + 1] = t.y * 1.4426950408889634f;   // scores in log2 units
    }
  } else {
#pragma unroll
    for (int d = 0; d < DQ; ++d) qr[d] = 0.f;
  }
  float o[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < HWk; k0 += KT) {
    const int nk = min(KT, HWk - k0);
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(kv + ((size_t)f * HWk + k0) * ROW);
    for (int i = threadIdx.x; i < nk * ROW / 8; i += 256) reinterpret_cast<uint4*>(&skv[0][0])[i] = __ldg(src + i);
    __syncthreads();
    // keys in blocks of 8: one running-max update / accumulator rescale per block instead of per key, and exp2 on
    // log2(e)-prescaled scores (softmax is shift-invariant, the scale folds into q): halves the MUFU work.
    for (int j0 = 0; j0 < nk; j0 += 8) {
      float sc[8];
      float bm = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        float s = 0.f;
        if (j0 + jj < nk) {
#pragma unroll
          for (int d = 0; d < DQ; d += 2) {
            const float2 t = unpack_h16x2(*reinterpret_cast<const uint32_t*>(&skv[j0 + jj][d]), fp16);
            s = fmaf(qr[d], t.x, s);
            s = fmaf(qr[d + 1], t.y, s);
          }
        } else {
          s = -INFINITY;
        }
        sc[jj] = s;
        bm = fmaxf(bm, s);
      }
      const float mn = fmaxf(m, bm);
      const float corr = ex2_approx(m - mn);          // m = -inf on the first block -> 0
      m = mn;
      l *= corr;
#pragma unroll
      for (int d = 0; d < 32; ++d) o[d] *= corr;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        if (j0 + jj < nk) {
          const float p = ex2_approx(sc[jj] - mn);
          l += p;
#pragma unroll
          for (int d = 0; d < 32; d += 2) {
            const float2 t = unpack_h16x2(*reinterpret_cast<const uint32_t*>(&skv[j0 + jj][DQ + slice * 32 + d]), fp16);
            o[d] = fmaf(p, t.x, o[d]);
            o[d + 1] = fmaf(p, t.y, o[d + 1]);
          }
        }
      }
    }
  }
  if (active) {
    const float inv = 1.0f / l;
    uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)f * HW + qi) * DV + slice * 32);
#pragma unroll
    for (int t = 0; t < 4; ++t)
      dst[t] = make_uint4(pack_h16x2(o[8 * t] * inv, o[8 * t + 1] * inv, fp16), pack_h16x2(o[8 * t + 2] * inv, o[8 * t + 3] * inv, fp16),
                          pack_h16x2(o[8 * t + 4] * inv, o[8 * t + 5] * inv, fp16), pack_h16x2(o[8 * t + 6] * inv, o[8 * t + 7] * inv, fp16));
  }
}

extern "C" int npvp_nonlocal_attention(const void* q, int64_t ldq, const void* kv, void* out, int64_t frames, int HW, int HWk, int dq,
                                       int dv, int fp16, void* stream) {
  NPVP_REQUIRE(q && kv && out && frames > 0 && HW > 0 && HWk > 0, "npvp_nonlocal_attention: bad arguments");
  NPVP_REQUIRE(dv == 4 * dq, "npvp_nonlocal_attention: expects dv = 4*dq (C/2 and C/8)");
  NPVP_REQUIRE(ldq % 2 == 0 && (uintptr_t)q % 4 == 0 && (uintptr_t)kv % 16 == 0 && (uintptr_t)out % 16 == 0, "npvp_nonlocal_attention: alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = dv / 32, QB = 256 / (S > 0 ? S : 1);
  const int64_t blocks = frames * ((HW + QB - 1) / QB);
  const h16* qq = (const h16*)q;
  const h16* kk = (const h16*)kv;
  h16* oo = (h16*)out;
  switch (dq) {
    case 8: nonlocal_attention_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(qq, ldq, kk, oo, HW, HWk, fp16); break;
    case 16: nonlocal_attention_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(qq, ldq, kk, oo, HW, HWk, fp16); break;
    case 32: nonlocal_attention_kernel<32><<<(unsigned)blocks, 256, 0, st>>>(qq, ldq, kk, oo, HW, HWk, fp16); break;
    case 64: nonlocal_attention_kernel<64><<<(unsigned)blocks, 256, 0, st>>>(qq, ldq, kk, oo, HW, HWk, fp16); break;
    default: NPVP_REQUIRE(false, "npvp_nonlocal_attention: dq must be 8, 16, 32 or 64 (got %d)", dq);
  }
  NPVP_LAUNCH_CHECK("nonlocal_attention_kernel");
  return NPVP_OK;
}
