#!/bin/bash
# builds tools/ubench/gemm_trace (diagnostic timeline of the GEMM pipeline; not part of the library)
cd "$(dirname "$0")/../.." && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Iinclude \
  tools/ubench/gemm_trace.cu npvp_b200/csrc/api.cu -o tools/ubench/gemm_trace -lcuda
