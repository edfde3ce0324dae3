// Micro-benchmark: cycles per tcgen05.mma (kind::f16, cta_group::1, both operands in shared memory) as a function of M, N and
// the swizzle mode of the operands.  One CTA per SM issues `reps` back-to-back MMAs (K = 16 each) and waits for the commit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -o tools/ubench/umma_shapes tools/ubench/umma_shapes.cu npvp_b200/csrc/api.cu
// Why: the tcgen05 head convolution (npvp_b200/csrc/head_tc.cu) first used M = 128, N = 32 MMAs from 64-byte-swizzled rows.
#include "../../npvp_b200/csrc/tc_ptx.cuh"
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint64_t desc(uint32_t addr, int rowb) {   // rowb = bytes per row = swizzle span (32 / 64 / 128)
  uint64_t d = (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((rowb * 8) >> 4) << 32) | ((uint64_t)1 << 46);
  return d | ((uint64_t)(rowb == 128 ? 2 : (rowb == 64 ? 4 : 6)) << 61);
}
__global__ void __launch_bounds__(128) k(long long* out, int M, int N, int rowb, int reps, int chain, int arows) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((uint32_t*)sm)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  ptx::fence_proxy_async();
  if (threadIdx.x < 32) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a0 = ptx::smem_u32(sm), b0 = a0 + 96 * 1024;
    const int ksteps = rowb / 32;
    // operand descriptors and accumulator addresses are precomputed: the issuing thread executes nothing but the MMAs
    // (a first version computed `i % arows` per MMA and measured its own integer divisions: 244 clk per MMA for every shape)
    uint64_t da[8], db[8];
    uint32_t dd[8];
    for (int i = 0; i < 8; ++i) {
      da[i] = desc(a0 + (uint32_t)(i % arows) * 8 * rowb, rowb) + 2 * (i % ksteps);     // A walks over row windows (the taps of the head convolution)
      db[i] = desc(b0 + (uint32_t)(i % 7) * 256 * rowb / 8, rowb) + 2 * (i % ksteps);
      dd[i] = slot + (chain ? 0 : (uint32_t)(i % (512 / N)) * N);
    }
    for (int pass = 0; pass < 2; ++pass) {
      const long long t0 = clock64();
      for (int i = 0; i < reps; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) ptx::umma_bf16(dd[j], da[j], db[j], idesc, chain ? (i + j > 0) : 0);
      }
      ptx::umma_commit(&bar);
      ptx::mbar_wait(&bar, pass & 1);
      const long long t1 = clock64();
      if (pass == 1) out[blockIdx.x] = t1 - t0;
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(slot, 512);
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 256;
  printf("cycles per tcgen05.mma (K = 16), %d back-to-back, 148 CTAs\n  M    N  rowB chain   clk/MMA   MAC/clk/SM\n", reps);
  for (int rowb : {128, 64}) for (int M : {128, 64}) for (int N : {256, 128, 64, 48, 32, 16}) for (int chain : {1, 0}) {
    if (M == 128 && N % 16) continue;
    k<<<148, 128, 200 * 1024>>>(d, M, N, rowb, reps, chain, 16);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error M %d N %d: %s\n", M, N, cudaGetErrorString(cudaGetLastError())); return 1; }
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double c = (double)mx / reps;
    printf("%4d %4d %5d %5d %9.1f %12.0f\n", M, N, rowb, chain, c, (double)M * N * 16 / c);
  }
  return 0;
}
