// Timeline of the persistent tcgen05 GEMM on one CTA (diagnostic; build: see tools/ubench/build_trace.sh).
//   ./gemm_trace M N K [conv: frames H W C KH]   -> per-tile / per-k-block hand-off times of CTA 0 in clocks
#define NPVP_GEMM_TRACE 1
#include "../../npvp_b200/csrc/gemm.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>

void npvp_set_ffn_scalar(int) {}
void npvp_set_ffn_mid16_mode(int) {}

static void dump(const char* name, int slot, int maxn) {
  static long long h[8][8192];
  int n[8];
  cudaMemcpyFromSymbol(h, g_trace, sizeof(h));
  cudaMemcpyFromSymbol(n, g_trace_n, sizeof(n));
  printf("%s (%d records):", name, n[slot]);
  long long t0 = h[3][0];
  for (int i = 0; i < n[slot] && i < maxn; ++i) printf(" %lld", h[slot][i] - t0);
  printf("\n");
}

int main(int argc, char** argv) {
  int64_t M = argc > 1 ? atoll(argv[1]) : 40960, N = argc > 2 ? atoll(argv[2]) : 512, K = argc > 3 ? atoll(argv[3]) : 512;
  const bool conv = argc > 4;
  int64_t frames = 0; int H = 0, W = 0, C = 0, KH = 0;
  if (conv) { frames = atoll(argv[4]); H = atoi(argv[5]); W = atoi(argv[6]); C = atoi(argv[7]); KH = atoi(argv[8]); M = frames * H * W; K = (int64_t)KH * KH * C; }
  if (getenv("EPI_DIRECT")) npvp_set_option("gemm_epi_direct", atoi(getenv("EPI_DIRECT")));
  if (getenv("GEMM_2CTA")) npvp_set_option("gemm_2cta", atoi(getenv("GEMM_2CTA")));
  void *A, *Wt, *out, *flush;
  cudaMalloc(&A, (conv ? frames * H * W * C : M * K) * 2);
  cudaMalloc(&Wt, N * K * 2);
  cudaMalloc(&out, M * N * 2);
  cudaMalloc(&flush, 256 << 20);
  cudaMemset(A, 0, (conv ? frames * H * W * C : M * K) * 2);
  cudaMemset(Wt, 0, N * K * 2);
  npvp_epilogue_t ep = {};
  ep.out_bf16 = out; ep.alpha = 1.f; ep.ld_out = N; ep.fp16 = 1;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(flush, rep, 256 << 20);
    int zero[8] = {0};
    cudaMemcpyToSymbol(g_trace_n, zero, sizeof(zero));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    int rc = conv ? npvp_conv_gemm_bf16(A, frames, H, W, C, KH, KH, 1, KH / 2, 0, H, W, 0, Wt, K, N, &ep, nullptr)
                  : npvp_gemm_bf16(A, K, Wt, K, M, N, K, &ep, 0, nullptr);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("rc %d (%s) cuda %s  M=%lld N=%lld K=%lld  %.1f us  %.1f TFLOP/s\n", rc, rc ? npvp_last_error() : "ok", cudaGetErrorString(err), (long long)M, (long long)N,
           (long long)K, ms * 1e3, 2.0 * M * N * K / ms / 1e9);
  }
  dump("producer: stage free", 0, 80);
  dump("mma: operands landed", 1, 80);
  dump("mma: k-block issued ", 2, 80);
  dump("mma: accumulator free (tile start)", 3, 20);
  dump("epilogue: accumulator full", 4, 20);
  dump("epilogue: tile stored     ", 5, 20);
  return 0;
}
