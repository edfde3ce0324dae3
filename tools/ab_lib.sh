#!/bin/bash
# same-box A/B of two builds of the library over the whole bench step: tools/ab_lib.sh npvp_b200/libnpvp_b200_prev.so   (prev, current, prev, current)
prev=$(realpath "$1")
for v in "$prev" "" "$prev" ""; do NPVP_B200_LIB=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('${v:-current}'.split('/')[-1], round(d['value']), 'frames/s', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], 'MHz  gemm', round(d['roofline']['kernel_ms_per_step'],2), 'ms', round(d['roofline']['frac'],3))
"; done
