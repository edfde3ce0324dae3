#!/bin/bash
# 2-GPU weak-scaling A/B of NCCL transport settings for the per-block gather (run on a 2-GPU box): tools/ab_nccl.sh
run() {
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$*', round(d['value']), 'frames/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value']), d['config'].get('gather_equal'))
"
}
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('1 GPU', round(d['value']), 'frames/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value']))
"
run A=0
run NCCL_P2P_USE_CUDA_MEMCPY=1
run NCCL_MAX_CTAS=1
run NCCL_MAX_CTAS=1 NCCL_P2P_USE_CUDA_MEMCPY=1
run A=0
