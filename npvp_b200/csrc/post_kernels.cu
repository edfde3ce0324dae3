// Pixel-space post-processing and evaluation metrics on the device: the consumers of the
// predicted frames.  Everything here is HBM-bound elementwise / stencil work on fp32 NCHW frames.
//   frames_to_pixels : VidReNormalize + clamp (+ uint8 as ToPILImage writes it)   utils/dataset.py:860-886, train_summary.py:243-248
//   pixels_to_frames : VidToTensor + VidNormalize                                 utils/dataset.py:835-858
//   psnr / ssim      : utils/metrics.py:12-30, 47-109 (per image)
// The arithmetic follows the reference operation by operation (IEEE div / sub, no FMA contraction) so the fp32 and uint8
// outputs are bit-identical to the reference transforms.
#include "common.cuh"

constexpr int kMaxC = 4;
struct ChanParams { float a[kMaxC], b[kMaxC]; };

// v = (x / inv_std[c]) - inv_mean[c], clamp [0,1]; u8 = trunc(v * 255)
__global__ void frames_to_pixels_kernel(const float* __restrict__ x, ChanParams p, float* __restrict__ out_f32, uint8_t* __restrict__ out_u8,
                                        int64_t total, int C, int64_t HW) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / HW) % C);
    float v = __fsub_rn(__fdiv_rn(__ldg(x + i), p.a[c]), p.b[c]);
    v = fminf(fmaxf(v, 0.0f), 1.0f);
    if (out_f32) out_f32[i] = v;
    if (out_u8) out_u8[i] = (uint8_t)__float2uint_rz(__fmul_rn(v, 255.0f));
  }
}

// 4 pixels per thread (HW % 4 == 0): one float4 in, one float4 and / or one 32-bit word of 4 bytes out
__global__ void frames_to_pixels_vec4_kernel(const float4* __restrict__ x, ChanParams p, float4* __restrict__ out_f32, uint32_t* __restrict__ out_u8,
                                             int64_t total4, int C, int64_t HW4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / HW4) % C);
    const float a = p.a[c], b = p.b[c];
    const float4 q = __ldg(x + i);
    float4 v;
    v.x = fminf(fmaxf(__fsub_rn(__fdiv_rn(q.x, a), b), 0.0f), 1.0f);
    v.y = fminf(fmaxf(__fsub_rn(__fdiv_rn(q.y, a), b), 0.0f), 1.0f);
    v.z = fminf(fmaxf(__fsub_rn(__fdiv_rn(q.z, a), b), 0.0f), 1.0f);
    v.w = fminf(fmaxf(__fsub_rn(__fdiv_rn(q.w, a), b), 0.0f), 1.0f);
    if (out_f32) out_f32[i] = v;
    if (out_u8)
      out_u8[i] = __float2uint_rz(__fmul_rn(v.x, 255.0f)) | (__float2uint_rz(__fmul_rn(v.y, 255.0f)) << 8) |
                  (__float2uint_rz(__fmul_rn(v.z, 255.0f)) << 16) | (__float2uint_rz(__fmul_rn(v.w, 255.0f)) << 24);
  }
}

extern "C" int npvp_frames_to_pixels(const float* frames, const float* inv_std, const float* inv_mean, float* out_f32, void* out_u8,
                                     int64_t n_images, int C, int64_t HW, void* stream) {
  NPVP_REQUIRE(frames && inv_std && inv_mean && (out_f32 || out_u8) && n_images > 0 && HW > 0, "npvp_frames_to_pixels: bad arguments");
  NPVP_REQUIRE(C >= 1 && C <= kMaxC, "npvp_frames_to_pixels: C must be in [1, %d] (got %d)", kMaxC, C);
  ChanParams p;
  for (int c = 0; c < C; ++c) { p.a[c] = inv_std[c]; p.b[c] = inv_mean[c]; }     // host pointers: a handful of scalars
  const int64_t total = n_images * C * HW;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = HW % 4 == 0 && (uintptr_t)frames % 16 == 0 && (!out_f32 || (uintptr_t)out_f32 % 16 == 0) && (!out_u8 || (uintptr_t)out_u8 % 4 == 0);
  if (vec) {
    const int64_t blocks = ceil_div64(total / 4, 256);
    frames_to_pixels_vec4_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>((const float4*)frames, p, (float4*)out_f32,
                                                                                                   (uint32_t*)out_u8, total / 4, C, HW / 4);
  } else {
    const int64_t blocks = ceil_div64(total, 256);
    frames_to_pixels_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(frames, p, out_f32, (uint8_t*)out_u8, total, C, HW);
  }
  NPVP_LAUNCH_CHECK("frames_to_pixels_kernel");
  return NPVP_OK;
}

// x = ((u8 / 255) - mean[c]) / std[c]
__global__ void pixels_to_frames_kernel(const uint8_t* __restrict__ in, ChanParams p, float* __restrict__ out, int64_t total, int C, int64_t HW) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / HW) % C);
    out[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)__ldg(in + i), 255.0f), p.a[c]), p.b[c]);
  }
}

extern "C" int npvp_pixels_to_frames(const void* in_u8, const float* mean, const float* std, float* out, int64_t n_images, int C, int64_t HW,
                                     void* stream) {
  NPVP_REQUIRE(in_u8 && mean && std && out && n_images > 0 && HW > 0, "npvp_pixels_to_frames: bad arguments");
  NPVP_REQUIRE(C >= 1 && C <= kMaxC, "npvp_pixels_to_frames: C must be in [1, %d] (got %d)", kMaxC, C);
  ChanParams p;
  for (int c = 0; c < C; ++c) { p.a[c] = mean[c]; p.b[c] = std[c]; }
  const int64_t total = n_images * C * HW;
  const int64_t blocks = ceil_div64(total, 256);
  pixels_to_frames_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)in_u8, p, out, total, C, HW);
  NPVP_LAUNCH_CHECK("pixels_to_frames_kernel");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// PSNR per image: -10 log10(mean((x/r - y/r)^2) + 1e-8); one block per image, fp32 per-thread partials, fp64 combine
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
psnr_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, int64_t elems, float data_range, int K, int T) {
  __shared__ double red[8];
  // K > 0: x holds K samples per clip, [clip][sample][t], scored against ONE ground truth [clip][t] (best-of-K evaluation)
  const size_t yimg = K > 0 ? (size_t)(blockIdx.x / (unsigned)(K * T)) * T + blockIdx.x % (unsigned)T : (size_t)blockIdx.x;
  const float* xi = x + (size_t)blockIdx.x * elems;
  const float* yi = y + yimg * elems;
  double acc = 0.0;
  for (int64_t i0 = 0; i0 < elems; i0 += 256 * 64) {           // fp32 within a 64-element run per thread, fp64 across runs
    float s = 0.f;
    for (int64_t i = i0 + threadIdx.x; i < elems && i < i0 + 256 * 64; i += 256) {
      const float d = __fdiv_rn(__ldg(xi + i), data_range) - __fdiv_rn(__ldg(yi + i), data_range);
      s = fmaf(d, d, s);
    }
    acc += (double)s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    const float mse = (float)(t / (double)elems);
    out[blockIdx.x] = -10.0f * log10f(mse + 1e-8f);
  }
}

extern "C" int npvp_psnr(const float* x, const float* y, float* out, int64_t n_images, int64_t elems, float data_range, void* stream) {
  NPVP_REQUIRE(x && y && out && n_images > 0 && elems > 0 && data_range > 0.f, "npvp_psnr: bad arguments");
  NPVP_REQUIRE(n_images < (1ll << 31), "npvp_psnr: too many images");
  psnr_kernel<<<(unsigned)n_images, 256, 0, (cudaStream_t)stream>>>(x, y, out, elems, data_range, 0, 1);
  NPVP_LAUNCH_CHECK("psnr_kernel");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// SSIM per image: 11x11 Gaussian (sigma 1.5), zero padding 5, depthwise; mean of the SSIM map over (C, H, W).
// One block per image walks its channels in 32x32 output tiles: the 42x42 input patches of both images are staged in
// shared memory, the five windowed sums (x, y, xx, yy, xy) are computed separably (11 horizontal + 11 vertical taps
// instead of 121), and the per-thread partial sums of the map are combined in fp64 - no atomics, so the result is
// deterministic.
// ---------------------------------------------------------------------------------------------
constexpr int kSsimW = 11, kSsimR = 5, kSsimT = 32, kSsimP = kSsimT + 2 * kSsimR;   // window, radius, tile, padded tile
struct SsimWin { float g[kSsimW]; };

__global__ void __launch_bounds__(256)
ssim_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, SsimWin win, int C, int H, int W, int K, int T) {
  __shared__ float px[kSsimP][kSsimP + 1], py[kSsimP][kSsimP + 1];
  __shared__ float hb[5][kSsimP][kSsimT + 1];
  __shared__ double red[8];
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  const size_t img = (size_t)blockIdx.x * C * H * W;
  const size_t yimg = (K > 0 ? (size_t)(blockIdx.x / (unsigned)(K * T)) * T + blockIdx.x % (unsigned)T : (size_t)blockIdx.x) * C * H * W;
  double acc = 0.0;
  for (int c = 0; c < C; ++c) {
    const float* xc = x + img + (size_t)c * H * W;
    const float* yc = y + yimg + (size_t)c * H * W;
    for (int ty = 0; ty < H; ty += kSsimT)
      for (int tx = 0; tx < W; tx += kSsimT) {
        __syncthreads();                                          // previous tile fully consumed
        for (int i = threadIdx.x; i < kSsimP * kSsimP; i += 256) {
          const int r = i / kSsimP, q = i % kSsimP;
          const int gy = ty + r - kSsimR, gx = tx + q - kSsimR;
          const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
          px[r][q] = in ? __ldg(xc + (size_t)gy * W + gx) : 0.f;
          py[r][q] = in ? __ldg(yc + (size_t)gy * W + gx) : 0.f;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kSsimP * kSsimT; i += 256) {   // horizontal pass
          const int r = i / kSsimT, q = i % kSsimT;
          float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
          for (int k = 0; k < kSsimW; ++k) {
            const float a = px[r][q + k], b = py[r][q + k], g = win.g[k];
            sx = fmaf(g, a, sx); sy = fmaf(g, b, sy);
            sxx = fmaf(g, a * a, sxx); syy = fmaf(g, b * b, syy); sxy = fmaf(g, a * b, sxy);
          }
          hb[0][r][q] = sx; hb[1][r][q] = sy; hb[2][r][q] = sxx; hb[3][r][q] = syy; hb[4][r][q] = sxy;
        }
        __syncthreads();
        float part = 0.f;
        for (int i = threadIdx.x; i < kSsimT * kSsimT; i += 256) {   // vertical pass + SSIM map
          const int r = i / kSsimT, q = i % kSsimT;
          if (ty + r >= H || tx + q >= W) continue;
          float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
          for (int k = 0; k < kSsimW; ++k) {
            const float g = win.g[k];
            m1 = fmaf(g, hb[0][r + k][q], m1); m2 = fmaf(g, hb[1][r + k][q], m2);
            e11 = fmaf(g, hb[2][r + k][q], e11); e22 = fmaf(g, hb[3][r + k][q], e22); e12 = fmaf(g, hb[4][r + k][q], e12);
          }
          const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
          const float s1 = e11 - m11, s2 = e22 - m22, s12 = e12 - m12;
          part += ((2.f * m12 + C1) * (2.f * s12 + C2)) / ((m11 + m22 + C1) * (s1 + s2 + C2));
        }
        acc += (double)part;
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[blockIdx.x] = (float)(t / ((double)C * H * W));
  }
}

extern "C" int npvp_ssim(const float* x, const float* y, const float* window11, float* out, int64_t n_images, int C, int H, int W, void* stream) {
  NPVP_REQUIRE(x && y && window11 && out && n_images > 0 && C > 0 && H > 0 && W > 0, "npvp_ssim: bad arguments");
  NPVP_REQUIRE(n_images < (1ll << 31), "npvp_ssim: too many images");
  SsimWin w;
  for (int k = 0; k < kSsimW; ++k) w.g[k] = window11[k];        // host pointer: the normalised 1-D Gaussian
  ssim_kernel<<<(unsigned)n_images, 256, 0, (cudaStream_t)stream>>>(x, y, out, w, C, H, W, 0, 1);
  NPVP_LAUNCH_CHECK("ssim_kernel");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// Best-of-K evaluation of stochastic samples (NPVP-S, BASELINE config 3: 8 samples per clip): per-frame PSNR or SSIM of
// every sample [clip][sample][t] against the clip's ground truth [clip][t] (read in place, never replicated), then per
// clip the sample with the best mean score over time is selected and its frames are copied out.
// ---------------------------------------------------------------------------------------------
extern "C" int npvp_sample_scores(const float* samples, const float* gt, const float* window11, float* scores, int64_t n_clips, int K, int T,
                                  int C, int H, int W, float data_range, void* stream) {
  NPVP_REQUIRE(samples && gt && scores && n_clips > 0 && K > 0 && T > 0 && C > 0 && H > 0 && W > 0 && data_range > 0.f, "npvp_sample_scores: bad arguments");
  const int64_t n_images = n_clips * K * T;
  NPVP_REQUIRE(n_images < (1ll << 31), "npvp_sample_scores: too many images");
  if (window11) {                                                  // SSIM (host pointer to the normalised 1-D Gaussian)
    SsimWin w;
    for (int k = 0; k < kSsimW; ++k) w.g[k] = window11[k];
    ssim_kernel<<<(unsigned)n_images, 256, 0, (cudaStream_t)stream>>>(samples, gt, scores, w, C, H, W, K, T);
    NPVP_LAUNCH_CHECK("ssim_kernel<samples>");
  } else {
    psnr_kernel<<<(unsigned)n_images, 256, 0, (cudaStream_t)stream>>>(samples, gt, scores, (int64_t)C * H * W, data_range, K, T);
    NPVP_LAUNCH_CHECK("psnr_kernel<samples>");
  }
  return NPVP_OK;
}

// one block per clip: mean score over time per sample (fp64), argmax (ties: lowest index), copy of the winner's frames
__global__ void __launch_bounds__(256)
best_of_k_kernel(const float* __restrict__ scores, const float* __restrict__ samples, int K, int T, int64_t clip_elems, int32_t* __restrict__ best_idx,
                 float* __restrict__ mean_scores, float* __restrict__ best) {
  __shared__ int s_best;
  const int n = blockIdx.x;
  if (threadIdx.x == 0) {
    int arg = 0;
    double top = 0.0;
    for (int k = 0; k < K; ++k) {
      double m = 0.0;
      for (int t = 0; t < T; ++t) m += (double)scores[((size_t)n * K + k) * T + t];
      m /= (double)T;
      mean_scores[(size_t)n * K + k] = (float)m;
      if (k == 0 || m > top) { top = m; arg = k; }
    }
    best_idx[n] = arg;
    s_best = arg;
  }
  __syncthreads();
  if (!best) return;
  const float4* src = reinterpret_cast<const float4*>(samples + ((size_t)n * K + s_best) * clip_elems);
  float4* dst = reinterpret_cast<float4*>(best + (size_t)n * clip_elems);
  for (int64_t i = threadIdx.x; i < clip_elems / 4; i += 256) dst[i] = __ldg(src + i);
}

extern "C" int npvp_best_of_k(const float* scores, const float* samples, int64_t n_clips, int K, int T, int64_t clip_elems, int32_t* best_idx,
                              float* mean_scores, float* best, void* stream) {
  NPVP_REQUIRE(scores && best_idx && mean_scores && n_clips > 0 && K > 0 && T > 0, "npvp_best_of_k: bad arguments");
  NPVP_REQUIRE(!best || (samples && clip_elems > 0 && clip_elems % 4 == 0 && ((uintptr_t)samples % 16) == 0 && ((uintptr_t)best % 16) == 0),
               "npvp_best_of_k: samples / best must be 16-byte aligned with clip_elems %% 4 == 0");
  NPVP_REQUIRE(n_clips < (1ll << 31), "npvp_best_of_k: too many clips");
  best_of_k_kernel<<<(unsigned)n_clips, 256, 0, (cudaStream_t)stream>>>(scores, samples, K, T, clip_elems, best_idx, mean_scores, best);
  NPVP_LAUNCH_CHECK("best_of_k_kernel");
  return NPVP_OK;
}
