for v in 0 1 0 1; do NPVP_B200_HEAD_TC=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('head_tc=$v', round(d['value']), d['ms_per_step'], d['clocks'])
"; done
