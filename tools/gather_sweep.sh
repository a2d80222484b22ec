#!/bin/bash
# Times the (K,H) = (8,8) gather kernels under each ring geometry of HAN_GATHER_CFG (attn_stream.cu) on the
# 2M-node bench workload; prints ms per step of the two kernels per setting.
out=${1:-gpurun_out/gather_sweep.txt}
: > "$out"
for cfg in 0,0 1,1 2,2 3,3; do
  HAN_GATHER_CFG=$cfg python bench.py --steps 5 --warmup 3 --no-parity --no-secondary --no-e2e --no-cpu-baseline 2>/dev/null |
    python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline'].get('kernels_ms_per_step',{})
print('$cfg', d['ms_per_step'], {n:v for n,v in k.items() if 'attn_fwd' in n or 'bwd_src' in n})" >> "$out"
done
cat "$out"
