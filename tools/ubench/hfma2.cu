// Micro-benchmark: packed half HFMA2 (fma.rn.f16x2) and MUFU.EX2.F16 issue throughput on sm_100a, next to scalar FFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/hfma2 tools/ubench/hfma2.cu
// The conv-FFN middle kernel (npvp_b200/csrc/ffn_mid16.cu) is ~55% HFMA2 and ~7% MUFU by instruction count: these rates are its
// arithmetic ceiling.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(unsigned* out, unsigned a, unsigned b, int iters) {
  unsigned x[16], y[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = 0x3c003c00u + threadIdx.x + i; y[i] = a + i; }
  if (MODE == 0) {            // HFMA2, accumulate in place, two shared operands
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
  } else if (MODE == 1) {     // HFMA2, three distinct register operands
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(y[(i + 5) & 15]));
    }
  } else if (MODE == 2) {     // MUFU.EX2.F16 x 2 + PRMT (ex2.approx.f16x2)
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
    }
  } else if (MODE == 3) {     // the kernel's mix: 6 HFMA2 : 1 ex2.f16x2 (2 MUFU + PRMT) : 1 max
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        unsigned p;
        asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(x[i]), "r"(a), "r"(b));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(p) : "r"(x[i]), "r"(y[i]));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(p) : "r"(x[i]), "r"(b));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(p));
        asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(p), "r"(y[(i + 3) & 15]));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(y[i]));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[(i + 7) & 15]), "r"(b));
      }
    }
  }
  else if (MODE == 4 || MODE == 5 || MODE == 6) {   // HFMA2 interleaved with scalar FFMA (4: 1:1, 5: 2:1) or FFMA alone (6): do the two share one pipe?
    float fa = __uint_as_float(a), fb = __uint_as_float(b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE != 6) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
        if (MODE == 5) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[8 + i]) : "r"(__float_as_uint(fa)), "r"(__float_as_uint(fb)));
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] ^= y[i];
  }
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int iters, double ops_per_iter_per_thread) {
  unsigned* out;
  cudaMalloc(&out, 148 * 2 * 512 * sizeof(unsigned));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<148 * 2, 512>>>(out, 0x3bff3bffu, 0x14001400u, iters);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 512>>>(out, 0x3bff3bffu, 0x14001400u, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instr = 148.0 * 2 * 16 * ops_per_iter_per_thread * iters;            // 16 warps per block
  printf("%s: %.3f ms, %.1f G warp-instr/s chip-wide = %.2f per clk per SM sub-partition at 1.9 GHz, err=%s\n", name, ms,
         warp_instr / ms * 1e-6, warp_instr / ms * 1e-6 / (148 * 4 * 1.9), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0>("HFMA2 shared operands  ", 1 << 13, 16);
  run<1>("HFMA2 distinct operands", 1 << 13, 16);
  run<2>("ex2.f16x2 (2 MUFU+PRMT)", 1 << 12, 16 * 3);
  run<3>("mix 6 HFMA2 : 1 ex2.f16x2", 1 << 12, 16 * 9);
  run<6>("FFMA alone (8 chains)  ", 1 << 13, 8);
  run<4>("mix 1 HFMA2 : 1 FFMA   ", 1 << 13, 16);
  run<5>("mix 2 HFMA2 : 1 FFMA   ", 1 << 13, 24);
  return 0;
}
