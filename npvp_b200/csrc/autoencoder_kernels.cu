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
// stem: fp32 NCHW image -> 16-bit NHWC features, reflect-pad 3, 7x7 conv + folded BN + ReLU, Cin in {1, 3}.
// Warp-level tensor-core formulation (the first CUDA-core version took 767 us per 128 Cityscapes frames): the input tile
// is staged in smem as 16-bit pixels of 4 channels (Cin zero-padded), so for a fixed kernel row ky the 7 taps x 4 channels
// of output pixel x are the 28 CONTIGUOUS halves starting at pixel x - one A row of a (16 pixels) x (K = 32) matrix whose
// last 4 columns hit zero weights.  Per ky: two mma.sync.m16n8k16 k-steps; 14 k-steps for the whole 7x7xCin window.
// Each warp owns one output row of the 8 x 64 tile (4 groups of 16 pixels) and 32 output channels (4 n-tiles).
// ---------------------------------------------------------------------------------------------
constexpr int kStemTY = 8, kStemTX = 64, kStemTW = kStemTX + 8;     // staged row: 64 + 6 halo + 1 read-ahead pixel, padded to 72

__device__ __forceinline__ void mma_16816_stem(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, int fp16) {
  if (fp16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

struct StemNorm { float mean[3], std[3]; };

// (r02 ncu: 139 registers allowed ONE 256-thread block per SM - 12% of the warp slots, 28% of the issue slots; capped at 128)
template <int CIN>
__global__ void __launch_bounds__(256, 2)
conv7x7_stem_kernel(const float* __restrict__ x, const uint8_t* __restrict__ x_u8, StemNorm nrm, const float* __restrict__ w,
                    const float* __restrict__ shift, h16* __restrict__ out, int Cout, int H, int W, int fp16, int tiles_per_block) {
  __shared__ __align__(16) h16 tile[kStemTY + 6][kStemTW][4];        // 8 KB
  __shared__ __align__(16) h16 wb[7][2][4][8][16];                   // [ky][k-step][n-tile][n][k]  14 KB
  const int groups = Cout / 32;
  const int f = blockIdx.z / groups, cg = blockIdx.z % groups;       // 32 output channels per block
  const int tiles_x = (W + kStemTX - 1) / kStemTX, tiles_y = (H + kStemTY - 1) / kStemTY;
  const int tx0 = (blockIdx.x % tiles_x) * kStemTX;
  // weights fp32 [(ky,kx,ci), Cout] -> B fragments: k = kx*4 + ci (kx = 7 and ci >= CIN are zero).  Once per block: the repack
  // (28 strided loads per thread) was 40 % of the instructions when every 8 x 64 tile did it, so a block now walks
  // tiles_per_block vertically adjacent tiles (as many as keep >= 2 blocks per SM busy).
  for (int i = threadIdx.x; i < 7 * 2 * 4 * 8 * 16; i += 256) {
    const int k = i & 15, n = (i >> 4) & 7, nt = (i >> 7) & 3, ks = (i >> 9) & 1, ky = i >> 10;
    const int kk = ks * 16 + k, kx = kk >> 2, ci = kk & 3;
    const float v = (kx < 7 && ci < CIN) ? __ldg(w + ((size_t)(ky * 7 + kx) * CIN + ci) * Cout + cg * 32 + nt * 8 + n) : 0.f;
    (&wb[0][0][0][0][0])[i] = float_to_h16(v, fp16);
  }
  const int wrp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const int tyb0 = (blockIdx.x / tiles_x) * tiles_per_block;
  for (int tyb = tyb0; tyb < min(tyb0 + tiles_per_block, tiles_y); ++tyb) {
  const int ty0 = tyb * kStemTY;
  if (tyb != tyb0) __syncthreads();                                  // every warp is done with the previous tile
  for (int i = threadIdx.x; i < (kStemTY + 6) * kStemTW; i += 256) {
    const int r = i / kStemTW, c = i % kStemTW;
    const int yy = reflect_idx(ty0 + r - 3, H), xx = reflect_idx(min(tx0 + c - 3, W + 2), W);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const size_t off = (((size_t)f * CIN + ci) * H + yy) * W + xx;
      // uint8 pixels: VidToTensor + VidNormalize in the reference's operation order, ((u8 / 255) - mean) / std (post_kernels.cu)
      v[ci] = x_u8 ? __fdiv_rn(__fsub_rn(__fdiv_rn((float)__ldg(x_u8 + off), 255.0f), nrm.mean[ci]), nrm.std[ci]) : __ldg(x + off);
    }
    *reinterpret_cast<uint2*>(&tile[r][c][0]) = make_uint2(pack_h16x2(v[0], v[1], fp16), pack_h16x2(v[2], v[3], fp16));
  }
  __syncthreads();
  float acc[4][4][4];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { acc[g][nt][0] = acc[g][nt][1] = acc[g][nt][2] = acc[g][nt][3] = 0.f; }
#pragma unroll 1
  for (int ky = 0; ky < 7; ++ky) {
    const uint32_t* trow = reinterpret_cast<const uint32_t*>(&tile[wrp + ky][0][0]);   // 2 words per pixel
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t b[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const uint32_t* bp = reinterpret_cast<const uint32_t*>(&wb[ky][ks][nt][gid][0]);
        b[nt][0] = bp[tig];
        b[nt][1] = bp[4 + tig];
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        // A[m][kk] = tile[row][x0 + m][kk] with kk running across pixels: word index (x0 + m)*2 + kk/2
        const uint32_t* ap = trow + (g * 16 + gid) * 2 + ks * 8 + tig;
        uint32_t a[4];
        a[0] = ap[0];          // row gid,     k = 2 tig
        a[1] = ap[16];         // row gid + 8  (8 pixels further = 16 words)
        a[2] = ap[4];          // row gid,     k = 2 tig + 8
        a[3] = ap[20];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_16816_stem(acc[g][nt], a, b[nt][0], b[nt][1], fp16);
      }
    }
  }
  const int oy = ty0 + wrp;
  if (oy < H) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int ox = tx0 + g * 16 + gid + hrow * 8;
        if (ox < W) {
          h16* dst = out + (((size_t)f * H + oy) * W + ox) * Cout + cg * 32 + 2 * tig;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const int co = cg * 32 + nt * 8 + 2 * tig;
            const float v0 = fmaxf(acc[g][nt][2 * hrow] + __ldg(shift + co), 0.f);
            const float v1 = fmaxf(acc[g][nt][2 * hrow + 1] + __ldg(shift + co + 1), 0.f);
            *reinterpret_cast<uint32_t*>(dst + nt * 8) = pack_h16x2(v0, v1, fp16);
          }
        }
      }
  }
  }
}

// head_tc.cu: the tcgen05 form of the stem (im2col row blocks built once per input row in a shared-memory ring); 1 = geometry not covered
int npvp_conv7x7_stem_tc_try(const float* x, const uint8_t* x_u8, const float* mean, const float* stdv, const float* w, const float* shift, void* out,
                             int64_t frames, int Cin, int Cout, int H, int W, int fp16, cudaStream_t st);

extern "C" int npvp_conv7x7_stem(const float* x, const float* w, const float* shift, void* out_bf16, int64_t frames, int Cin,
                                 int Cout, int H, int W, int fp16, const void* x_u8, const float* norm_mean, const float* norm_std,
                                 void* stream) {
  NPVP_REQUIRE((x != nullptr) != (x_u8 != nullptr), "npvp_conv7x7_stem: exactly one of x (fp32 frames) and x_u8 (uint8 pixels) must be given");
  NPVP_REQUIRE(!x_u8 || (norm_mean && norm_std), "npvp_conv7x7_stem: uint8 input needs the VidNormalize mean / std (host pointers)");
  NPVP_REQUIRE(w && shift && out_bf16 && frames > 0, "npvp_conv7x7_stem: bad arguments");
  StemNorm nrm = {};
  if (x_u8) for (int c = 0; c < Cin && c < 3; ++c) { nrm.mean[c] = norm_mean[c]; nrm.std[c] = norm_std[c]; }
  NPVP_REQUIRE(Cout % 32 == 0 && H >= 4 && W >= 4, "npvp_conv7x7_stem: Cout must be a multiple of 32, H/W >= 4");
  {
    const int rc = npvp_conv7x7_stem_tc_try(x, (const uint8_t*)x_u8, norm_mean, norm_std, w, shift, out_bf16, frames, Cin, Cout, H, W, fp16, (cudaStream_t)stream);
    if (rc <= 0) return rc;
  }
  NPVP_REQUIRE(frames * (Cout / 32) <= 65535, "npvp_conv7x7_stem: too many frames per launch (%lld)", (long long)frames);
  const int tiles_y = (H + kStemTY - 1) / kStemTY, tiles_x = (W + kStemTX - 1) / kStemTX;
  // vertically adjacent tiles per block: as many as still leave ~2 blocks per SM (296 on a B200) in flight
  int tpb = (int)((frames * (Cout / 32) * tiles_x * tiles_y) / 296);
  tpb = tpb < 1 ? 1 : (tpb > tiles_y ? tiles_y : tpb);
  tpb = (tiles_y + (tiles_y + tpb - 1) / tpb - 1) / ((tiles_y + tpb - 1) / tpb);     // equal shares for the blocks of a column
  dim3 grid((unsigned)(((tiles_y + tpb - 1) / tpb) * tiles_x), 1, (unsigned)(frames * (Cout / 32)));
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 1) conv7x7_stem_kernel<1><<<grid, 256, 0, st>>>(x, (const uint8_t*)x_u8, nrm, w, shift, (h16*)out_bf16, Cout, H, W, fp16, tpb);
  else if (Cin == 3) conv7x7_stem_kernel<3><<<grid, 256, 0, st>>>(x, (const uint8_t*)x_u8, nrm, w, shift, (h16*)out_bf16, Cout, H, W, fp16, tpb);
  else NPVP_REQUIRE(false, "npvp_conv7x7_stem: Cin must be 1 or 3 (got %d)", Cin);
  NPVP_LAUNCH_CHECK("conv7x7_stem_kernel");
  return NPVP_OK;
}

// ---------------------------------------------------------------------------------------------
// head: 16-bit NHWC (or phase-major) features -> fp32 NCHW image, reflect-pad 3, + bias + tanh/sigmoid.
// Cout is 1 or 3, so a 128-row UMMA tile would stream 3 KB of A per output pixel for almost no math, and a per-tap
// mma (N = Cout padded to 8) re-reads every A fragment 49 times from shared memory (the second version of this kernel:
// ldmatrix-bound, 1.93 ms per 640 Cityscapes frames).  This version folds the horizontal taps into N and the vertical
// taps into K:
//     Z[x', (kx, co)] = sum_{ky, ci} X[oy + ky, x', ci] * W[ky, kx, ci, co]        one GEMM per output row,
//                                                                                  M = staged pixels, N = 7 Cout (-> 8 NT),
//                                                                                  K = 7 x 32 channels per pass
//     out[oy, ox, co]  = bias[co] + sum_kx Z[ox + kx, (kx, co)]                     shift-sum through shared memory
// so each A fragment (16 pixels x 16 channels, one ldmatrix.x4) feeds NT mma.sync.m16n8k16 and is read once per ky
// instead of once per tap: 5.6x fewer ldmatrix and 1.9x fewer mma for Cout = 3.
// Block = 8 x 64 output tile, 8 warps (one output row each); the input tile (14 rows x 70 pixels, 32 channels per pass,
// 64-byte pixels with a 16-byte-chunk XOR swizzle so the 8 ldmatrix rows hit distinct banks) is staged with cp.async;
// Cin = 64 takes two passes over the same buffers.  The per-warp Z scratch aliases the tile after the last pass.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <bool FP16>
__device__ __forceinline__ void mma_16816_head(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (FP16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
__device__ __forceinline__ void head_cp_async_16(uint32_t smem_dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(src) : "memory");
}

constexpr int kHeadTY = 8, kHeadTX = 64;                       // output tile per block (8 warps: one output row each)
constexpr int kHeadRows = kHeadTY + 6;                         // staged input rows
constexpr int kHeadCols = kHeadTX + 6;                         // staged input pixels per row that feed valid outputs
constexpr int kHeadGroups = (kHeadCols + 15) / 16;             // 16-pixel mma row groups per output row (5)
constexpr int kHeadTW = kHeadGroups * 16;                      // smem row pitch in pixels (80; the last 10 are never staged)
constexpr int kHeadTileBytes = kHeadRows * kHeadTW * 64;       // 32 channels x 2 B per pixel

struct HeadPix { float inv_std[3], inv_mean[3]; };

template <int NT, bool FP16>
__global__ void __launch_bounds__(256, 2)
conv7x7_head_kernel(const h16* __restrict__ x, const uint2* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                    int Cin, int Cout, int H, int W, int phase_major, int act, uint8_t* __restrict__ out_u8, HeadPix pix) {
  constexpr int SST = NT * 8 + 1;                               // scratch row stride (floats), odd: conflict-free shift-sum
  static_assert(8 * kHeadTW * SST * 4 <= kHeadTileBytes, "Z scratch must fit in the tile it aliases");
  extern __shared__ __align__(128) uint8_t head_smem[];
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(head_smem);
  uint2* wb = reinterpret_cast<uint2*>(head_smem + kHeadTileBytes);    // [pass][14 k-steps][NT][32 lanes] B fragments
  const uint32_t wb_s = tile_s + kHeadTileBytes;
  const int passes = Cin / 32;
  const int f = blockIdx.z;
  const int tiles_x = (W + kHeadTX - 1) / kHeadTX;
  const int ty0 = (blockIdx.x / tiles_x) * kHeadTY, tx0 = (blockIdx.x % tiles_x) * kHeadTX;
  const int wrp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;

  for (int i = threadIdx.x; i < passes * 14 * NT * 16; i += 256)          // 16-byte vectors of the packed weights
    head_cp_async_16(wb_s + i * 16, reinterpret_cast<const uint4*>(w) + i);

  // per-lane ldmatrix.x4 source: pixel lm_px of the group, 16-byte chunk (2 cb + lm_hi) XOR-swizzled by the pixel index
  const int lm_px = ((lane >> 3) & 1) * 8 + (lane & 7), lm_hi = lane >> 4, lm_sw = (lm_px >> 1) & 3;
  float acc[kHeadGroups][NT][4];
#pragma unroll
  for (int g = 0; g < kHeadGroups; ++g)
#pragma unroll
    for (int j = 0; j < NT; ++j) { acc[g][j][0] = acc[g][j][1] = acc[g][j][2] = acc[g][j][3] = 0.f; }

  for (int pass = 0; pass < passes; ++pass) {
    if (pass) __syncthreads();                                  // every warp is done with the previous channel slice
    for (int i = threadIdx.x; i < kHeadRows * kHeadCols; i += 256) {
      const int r = i / kHeadCols, c = i % kHeadCols;
      const int yy = min(reflect_idx(ty0 + r - 3, H), H - 1), xx = min(reflect_idx(tx0 + c - 3, W), W - 1);
      const h16* src = x + pixel_offset(f, yy, xx, H, W, phase_major) * Cin + pass * 32;
      const uint32_t dst = tile_s + (uint32_t)(r * kHeadTW + c) * 64;
      const int sw = (c >> 1) & 3;
#pragma unroll
      for (int v = 0; v < 4; ++v) head_cp_async_16(dst + ((v ^ sw) << 4), src + v * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const uint2* wp = wb + (size_t)pass * 14 * NT * 32 + lane;
#pragma unroll
    for (int s = 0; s < 14; ++s) {                              // k-step = (ky, 16-channel half)
      const int ky = s >> 1, cb = s & 1;
      uint2 b[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) b[j] = wp[(s * NT + j) * 32];
      const uint32_t a_row = tile_s + (uint32_t)((wrp + ky) * kHeadTW + lm_px) * 64 + (uint32_t)(((cb * 2 + lm_hi) ^ lm_sw) << 4);
#pragma unroll
      for (int g = 0; g < kHeadGroups; ++g) {
        uint32_t a[4];
        ldsm_x4(a, a_row + g * 16 * 64);
#pragma unroll
        for (int j = 0; j < NT; ++j) mma_16816_head<FP16>(acc[g][j], a, b[j].x, b[j].y);
      }
    }
  }
  __syncthreads();                                              // the tile is dead: reuse it as per-warp Z scratch
  float* S = reinterpret_cast<float*>(head_smem) + (size_t)wrp * kHeadTW * SST;
#pragma unroll
  for (int g = 0; g < kHeadGroups; ++g)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) S[(g * 16 + gid + (e >> 1) * 8) * SST + j * 8 + 2 * tig + (e & 1)] = acc[g][j][e];
  __syncwarp();
  const int oy = ty0 + wrp;
  if (oy >= H) return;
  for (int i = lane; i < kHeadTX * Cout; i += 32) {
    const int co = i / kHeadTX, ox = i % kHeadTX;
    if (tx0 + ox >= W) continue;
    float v = __ldg(bias + co);
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) v += S[(ox + kx) * SST + kx * Cout + co];
    v = (act == NPVP_ACT_TANH) ? tanhf(v) : (act == NPVP_ACT_SIGMOID ? 1.0f / (1.0f + expf(-v)) : v);
    const size_t o = (((size_t)f * Cout + co) * H + oy) * W + tx0 + ox;
    if (out) out[o] = v;
    if (out_u8) {   // VidReNormalize + clamp + ToPILImage in the reference's operation order (frames_to_pixels_kernel): bit-identical
      const float p = fminf(fmaxf(__fsub_rn(__fdiv_rn(v, pix.inv_std[co]), pix.inv_mean[co]), 0.0f), 1.0f);
      out_u8[o] = (uint8_t)__float2uint_rz(__fmul_rn(p, 255.0f));
    }
  }
}

template <int NT, bool FP16>
static int launch_head(const void* x, const void* w, const float* bias, float* out, int64_t frames, int Cin, int Cout, int H, int W,
                       int phase_major, int act, uint8_t* out_u8, const HeadPix& pix, cudaStream_t st) {
  const int smem = kHeadTileBytes + (Cin / 32) * 14 * NT * 32 * 8;
  static int attr_smem = 0;
  if (attr_smem < smem) {
    cudaError_t err = cudaFuncSetAttribute(conv7x7_head_kernel<NT, FP16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) { npvp_set_error("conv7x7_head: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(err)); return NPVP_ERR_CUDA; }
    attr_smem = smem;
  }
  dim3 grid((unsigned)(((H + kHeadTY - 1) / kHeadTY) * ((W + kHeadTX - 1) / kHeadTX)), 1, (unsigned)frames);
  conv7x7_head_kernel<NT, FP16><<<grid, 256, smem, st>>>((const h16*)x, (const uint2*)w, bias, out, Cin, Cout, H, W, phase_major, act, out_u8, pix);
  NPVP_LAUNCH_CHECK("conv7x7_head_kernel");
  return NPVP_OK;
}

// head_tc.cu: the tcgen05 form (whole padded rows streamed through a shared-memory ring); 1 = geometry not covered
int npvp_conv7x7_head_tc_try(const void* x, const void* w, const float* bias, float* out, int64_t frames, int Cin, int Cout, int H, int W,
                             int act, int fp16, uint8_t* out_u8, const float* pix_inv_std, const float* pix_inv_mean, cudaStream_t st);

extern "C" int npvp_conv7x7_head(const void* x_bf16, const void* w, const float* bias, float* out, int64_t frames, int Cin,
                                 int Cout, int H, int W, int phase_major, int act, int fp16, void* out_u8, const float* pix_inv_std,
                                 const float* pix_inv_mean, void* stream) {
  NPVP_REQUIRE(x_bf16 && w && bias && (out || out_u8) && frames > 0 && frames <= 65535, "npvp_conv7x7_head: bad arguments");
  NPVP_REQUIRE(!out_u8 || (pix_inv_std && pix_inv_mean), "npvp_conv7x7_head: the uint8 output needs the VidReNormalize constants (host pointers)");
  HeadPix pix = {};
  if (out_u8) for (int c = 0; c < Cout && c < 3; ++c) { pix.inv_std[c] = pix_inv_std[c]; pix.inv_mean[c] = pix_inv_mean[c]; }
  NPVP_REQUIRE(H >= 4 && W >= 4 && (!phase_major || (H % 2 == 0 && W % 2 == 0)), "npvp_conv7x7_head: H/W >= 4, even H/W for phase-major input");
  NPVP_REQUIRE(Cout >= 1 && Cout <= 3, "npvp_conv7x7_head: Cout must be in [1, 3] (got %d)", Cout);
  NPVP_REQUIRE(Cin == 32 || Cin == 64, "npvp_conv7x7_head: Cin must be 32 or 64 (got %d)", Cin);
  NPVP_REQUIRE((uintptr_t)w % 16 == 0 && (uintptr_t)x_bf16 % 16 == 0, "npvp_conv7x7_head: x / w must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (!phase_major) {
    const int rc = npvp_conv7x7_head_tc_try(x_bf16, w, bias, out, frames, Cin, Cout, H, W, act, fp16, (uint8_t*)out_u8, pix_inv_std, pix_inv_mean, st);
    if (rc <= 0) return rc;
  }
  const int nt = (7 * Cout + 7) / 8;
#define NPVP_HEAD_CASE(NTV)                                                                                                   \
  if (nt == NTV)                                                                                                              \
    return fp16 ? launch_head<NTV, true>(x_bf16, w, bias, out, frames, Cin, Cout, H, W, phase_major, act, (uint8_t*)out_u8, pix, st)  \
                : launch_head<NTV, false>(x_bf16, w, bias, out, frames, Cin, Cout, H, W, phase_major, act, (uint8_t*)out_u8, pix, st);
  NPVP_HEAD_CASE(1)
  NPVP_HEAD_CASE(2)
  NPVP_HEAD_CASE(3)
#undef NPVP_HEAD_CASE
  return NPVP_ERR_INVALID;
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
// non-local attention core (flash style, UNSCALED softmax): one warp = 16 queries x one <=64-wide slice of dv.
// Keys/values stream through smem in tiles of 64; S = Q K^T and O += P V run on warp-level mma.sync.m16n8k16
// (fp16 or bf16 operands to match the activations, fp32 accumulate), the online softmax lives in the accumulator
// registers (row max / sum via quad shuffles, exp2 on log2(e)-scaled scores).  The first SIMT version was FMA-bound
// (40 FMA per query-key pair, 2.3 ms per launch at the 64x64 stage); the tensor path needs ~1/25 of the instructions.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, int fp16) {
  if (fp16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
// K = 8 form for dq = 8 (the 64 x 64 stage): half the tensor-pipe work of the zero-padded m16n8k16 - the kernel issues 24 HMMA per
// 64-key tile and warp there, and ncu shows math_pipe_throttle next to a MUFU pipe that is only 44 % busy
__device__ __forceinline__ void mma_1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0, int fp16) {
  if (fp16) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
  } else {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
  }
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

// r02: the pooled keys / values of a WHOLE frame stay resident in shared memory (HWk x 5 dq 16-bit values: 80 KB at the 64 x 64
// stage, less elsewhere), fetched once per block with cp.async in chunks of 256 keys that are waited for one by one, so the
// key loop has no block-wide barrier and no exposed load (the first version re-staged every 64-key tile through registers
// between two __syncthreads: 336 us per 128 frames at 64 x 64 against 117 us of MUFU time).  A block is 8 warps = 128 queries.
// (4-warp blocks where a frame has few queries: 8 x 8 frames would leave half of an 8-warp block idle - 24.7 vs 15.5 us)
constexpr int kNlChunk = 256;
template <int DQ, int kNlWarps>
__global__ void __launch_bounds__(kNlWarps * 32)
nonlocal_attention_kernel(const h16* __restrict__ q, int64_t ldq, const h16* __restrict__ kv, h16* __restrict__ out, int HW, int HWk, int fp16) {
  constexpr int DV = 4 * DQ, DVS = DV < 64 ? DV : 64, NT = DVS / 8, KT = 64, ROW = DQ + DV;
  constexpr int QK_STEPS = DQ < 16 ? 1 : DQ / 16;
  extern __shared__ __align__(16) uint8_t nl_smem[];
  h16 (*skv)[ROW] = reinterpret_cast<h16 (*)[ROW]>(nl_smem);          // [keys rounded up to 64][ROW]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  constexpr int QB = kNlWarps * 16;
  const int blocks_per_frame = (HW + QB - 1) / QB;
  const int64_t f = blockIdx.x / blocks_per_frame;
  const int q_base = (blockIdx.x % blocks_per_frame) * QB + w * 16;
  {
    // all keys / values of the frame: 16-byte cp.async, one commit group per 256-key chunk; rows past HWk are zero-filled
    const int keys_pad = (HWk + KT - 1) / KT * KT;
    const h16* src = kv + (size_t)f * HWk * ROW;
    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(nl_smem);
    constexpr int VEC_PER_KEY = ROW / 8;
    for (int c0 = 0; c0 < keys_pad; c0 += kNlChunk) {
      const int c1 = min(c0 + kNlChunk, keys_pad);
      for (int i = c0 * VEC_PER_KEY + (int)threadIdx.x; i < c1 * VEC_PER_KEY; i += kNlWarps * 32) {
        const bool live = i < HWk * VEC_PER_KEY;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (uint32_t)i * 16u), "l"(src + (size_t)(live ? i : 0) * 8), "r"(live ? 16u : 0u) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
  const int dv0 = blockIdx.y * DVS;                        // dv slice of this block
  const int r0 = q_base + gid, r1 = q_base + gid + 8;
  const bool ok0 = r0 < HW, ok1 = r1 < HW;
  // Q as A fragments (rows = queries, k = d); DQ = 8 pads k 8..15 with zeros
  uint32_t qa[QK_STEPS][4];
  {
    const uint32_t* q0 = reinterpret_cast<const uint32_t*>(q + ((size_t)f * HW + (ok0 ? r0 : 0)) * ldq);
    const uint32_t* q1 = reinterpret_cast<const uint32_t*>(q + ((size_t)f * HW + (ok1 ? r1 : 0)) * ldq);
#pragma unroll
    for (int ks = 0; ks < QK_STEPS; ++ks) {
      qa[ks][0] = ok0 ? __ldg(q0 + 8 * ks + tig) : 0u;
      qa[ks][1] = ok1 ? __ldg(q1 + 8 * ks + tig) : 0u;
      qa[ks][2] = (DQ >= 16 && ok0) ? __ldg(q0 + 8 * ks + 4 + tig) : 0u;
      qa[ks][3] = (DQ >= 16 && ok1) ? __ldg(q1 + 8 * ks + 4 + tig) : 0u;
    }
  }
  float o[NT][4];
#pragma unroll
  for (int d = 0; d < NT; ++d) { o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  constexpr float kLog2e = 1.4426950408889634f;

  const int n_chunks = (HWk + kNlChunk - 1) / kNlChunk;
  for (int k0 = 0; k0 < HWk; k0 += KT) {
    const int nk = min(KT, HWk - k0);
    if (k0 % kNlChunk == 0) {                                    // entering a new chunk: wait for its commit group only
      const int pending = n_chunks - 1 - k0 / kNlChunk;          // groups that may still be in flight afterwards
      if (pending >= 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
      else if (pending == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
      else if (pending == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    }
    // ---- S = Q K^T for 64 keys: 8 key tiles ----
    float sc[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      sc[t][0] = sc[t][1] = sc[t][2] = sc[t][3] = 0.f;
      const uint32_t* krow = reinterpret_cast<const uint32_t*>(&skv[k0 + 8 * t + gid][0]);
#pragma unroll
      for (int ks = 0; ks < QK_STEPS; ++ks) {
        const uint32_t b0 = krow[8 * ks + tig];
        if (DQ < 16) { mma_1688(sc[t], qa[ks][0], qa[ks][1], b0, fp16); continue; }
        const uint32_t b1 = krow[8 * ks + 4 + tig];
        mma_16816(sc[t], qa[ks], b0, b1, fp16);
      }
    }
    // ---- online softmax: row maxima on the raw scores, p = 2^(s log2e - m log2e) as ONE FFMA + ONE MUFU per element.
    // Only a ragged last key tile needs masking (r02 ncu at the 64x64 stage: 224 M warp instructions per launch, the
    // per-element FMUL + 2 selects + FSUB of the first version were a quarter of them).
    float bm0 = -INFINITY, bm1 = -INFINITY;
    if (nk < KT) {
#pragma unroll
      for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool dead = (8 * t + 2 * tig + e) >= nk;
          if (dead) { sc[t][e] = -INFINITY; sc[t][2 + e] = -INFINITY; }
        }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        bm0 = fmaxf(bm0, sc[t][e]);
        bm1 = fmaxf(bm1, sc[t][2 + e]);
      }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float mn0 = fmaxf(m0, bm0 * kLog2e), mn1 = fmaxf(m1, bm1 * kLog2e);   // running maxima in the log2 domain
    const float c0 = ex2_approx(m0 - mn0), c1 = ex2_approx(m1 - mn1);     // 0 on the first tile (m = -inf)
    m0 = mn0; m1 = mn1;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float p0 = ex2_approx(fmaf(sc[t][e], kLog2e, -mn0)), p1 = ex2_approx(fmaf(sc[t][2 + e], kLog2e, -mn1));   // ex2(-inf) = 0 for dead keys
        sc[t][e] = p0; sc[t][2 + e] = p1;
        s0 += p0; s1 += p1;
      }
    l0 = l0 * c0 + s0;       // per-thread partial row sums; quad-reduced once at the end
    l1 = l1 * c1 + s1;
#pragma unroll
    for (int d = 0; d < NT; ++d) { o[d][0] *= c0; o[d][1] *= c0; o[d][2] *= c1; o[d][3] *= c1; }
    // ---- O += P V: 4 steps of 16 keys ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      a[0] = pack_h16x2(sc[2 * kk][0], sc[2 * kk][1], fp16);
      a[1] = pack_h16x2(sc[2 * kk][2], sc[2 * kk][3], fp16);
      a[2] = pack_h16x2(sc[2 * kk + 1][0], sc[2 * kk + 1][1], fp16);
      a[3] = pack_h16x2(sc[2 * kk + 1][2], sc[2 * kk + 1][3], fp16);
#pragma unroll
      for (int d = 0; d < NT; ++d) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, &skv[k0 + 16 * kk + (lane & 15)][DQ + dv0 + 8 * d]);
        mma_16816(o[d], a, b0, b1, fp16);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
#pragma unroll
  for (int d = 0; d < NT; ++d) {
    if (ok0) reinterpret_cast<uint32_t*>(out + ((size_t)f * HW + r0) * DV + dv0)[4 * d + tig] = pack_h16x2(o[d][0] * inv0, o[d][1] * inv0, fp16);
    if (ok1) reinterpret_cast<uint32_t*>(out + ((size_t)f * HW + r1) * DV + dv0)[4 * d + tig] = pack_h16x2(o[d][2] * inv1, o[d][3] * inv1, fp16);
  }
}

extern "C" int npvp_nonlocal_attention(const void* q, int64_t ldq, const void* kv, void* out, int64_t frames, int HW, int HWk, int dq,
                                       int dv, int fp16, void* stream) {
  NPVP_REQUIRE(q && kv && out && frames > 0 && HW > 0 && HWk > 0, "npvp_nonlocal_attention: bad arguments");
  NPVP_REQUIRE(dv == 4 * dq, "npvp_nonlocal_attention: expects dv = 4*dq (C/2 and C/8)");
  NPVP_REQUIRE(ldq % 2 == 0 && (uintptr_t)q % 4 == 0 && (uintptr_t)kv % 16 == 0 && (uintptr_t)out % 16 == 0, "npvp_nonlocal_attention: alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int dvs = dv < 64 ? dv : 64;
  const int warps = HW >= 1024 ? 8 : 4;
  const int QB = warps * 16;
  const dim3 grid((unsigned)(frames * ((HW + QB - 1) / QB)), (unsigned)(dv / dvs));
  const h16* qq = (const h16*)q;
  const h16* kk = (const h16*)kv;
  h16* oo = (h16*)out;
  const int smem = (HWk + 63) / 64 * 64 * 5 * dq * 2;                 // every key / value row of a frame
  NPVP_REQUIRE(smem <= 200 * 1024, "npvp_nonlocal_attention: %d pooled keys x %d values do not fit in shared memory", HWk, 5 * dq);
#define NPVP_NL_W(DQ_, W_) { \
    static int attr = 0; \
    if (attr < smem) { \
      cudaError_t err = cudaFuncSetAttribute(nonlocal_attention_kernel<DQ_, W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
      if (err != cudaSuccess) { npvp_set_error("nonlocal_attention: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(err)); return NPVP_ERR_CUDA; } \
      attr = smem; \
    } \
    nonlocal_attention_kernel<DQ_, W_><<<grid, W_ * 32, smem, st>>>(qq, ldq, kk, oo, HW, HWk, fp16); }
#define NPVP_NL(DQ_) { if (warps == 8) NPVP_NL_W(DQ_, 8) else NPVP_NL_W(DQ_, 4) }
  switch (dq) {
    case 8: NPVP_NL(8); break;
    case 16: NPVP_NL(16); break;
    case 32: NPVP_NL(32); break;
    case 64: NPVP_NL(64); break;
    default: NPVP_REQUIRE(false, "npvp_nonlocal_attention: dq must be 8, 16, 32 or 64 (got %d)", dq);
  }
#undef NPVP_NL
#undef NPVP_NL_W
  NPVP_LAUNCH_CHECK("nonlocal_attention_kernel");
  return NPVP_OK;
}
