#!/bin/bash
# Round check on the GPU box: tests, smoke, bench, launch list, one full ncu capture of the dominant kernel.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/ci_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/ci_smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/ci_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/ci_launches.csv python bench.py --profile --clips 64 > gpurun_out/ci_ncu_launches.log 2>&1
# the largest GEMM family of the step: fc1 of the conv-FFN with the frame-statistics epilogue (M = 40960, N = 2048, K = 512)
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k regex:"v2_kernel<\(int\)256, \(int\)0, \(int\)0, \(int\)12, \(int\)0>" -s 8 -c 2 \
    -f -o gpurun_out/ci_gemm_full python bench.py --profile --clips 64 > gpurun_out/ci_ncu_full.log 2>&1
timeout 200 python tools/bench_configs.py --faithful > gpurun_out/ci_faithful.md 2>&1
