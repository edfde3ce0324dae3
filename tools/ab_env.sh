#!/bin/bash
# same-box A/B of one environment switch over the whole bench step: tools/ab_env.sh NPVP_B200_HEAD_TC   (runs 0 1 0 1)
var=$1
for v in 0 1 0 1; do env $var=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$var=$v', round(d['value']), 'frames/s', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], 'MHz', [(f['kernel'], round(f['ms_per_step'],2), round(f['frac'],2)) for f in d['roofline_memory'] if f['kernel'].startswith(('frame','conv7','add_ln'))])
"; done
