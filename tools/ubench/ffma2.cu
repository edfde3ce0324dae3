// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ffma2 tools/ubench/ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, float a, float b, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
    }
  } else if (MODE == 1) {
    unsigned long long p[8], aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(aa), "l"(bb));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(p[i]));
  }
  if (MODE == 2) {            // scalar, three distinct register operands per FMA (no operand reuse between neighbours)
    float y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = a + i * 1e-6f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(y[i]), "f"(y[(i + 5) & 15]));
    }
  }
  if (MODE == 3) {            // packed, three distinct 64-bit register operands per FFMA2
    unsigned long long p[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
      asm("mov.b64 %0, {%1, %2};" : "=l"(y[i]) : "f"(a + i * 1e-6f), "f"(a - i * 1e-6f));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(y[i]), "l"(y[(i + 3) & 7]));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(p[i]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int iters) {
  float* out;
  cudaMalloc(&out, 148 * 2 * 512 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<148 * 2, 512>>>(out, 0.999f, 0.001f, iters);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 512>>>(out, 0.999f, 0.001f, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double fma = 148.0 * 2 * 512 * 16.0 * iters;
  printf("%s: %.3f ms, %.2f TFMA/s (%.1f fp32 FMA lanes per ns chip-wide), err=%s\n", name, ms, fma / ms * 1e-9, fma / ms * 1e-6,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0>("FFMA  (scalar, 16 chains/thread)", 1 << 14);
  run<1>("FFMA2 (packed,  8 chains/thread)", 1 << 14);
  run<2>("FFMA  (scalar, distinct operands)", 1 << 14);
  run<3>("FFMA2 (packed, distinct operands)", 1 << 14);
  run<2>("FFMA  (scalar, distinct operands)", 1 << 14);
  run<3>("FFMA2 (packed, distinct operands)", 1 << 14);
  return 0;
}
