#!/usr/bin/env bash
# A/B of one environment variable over arbitrary values on ONE box, alternating runs.
#   gpurun -- 'bash tools/ab_values.sh DDP_B200_GEMM_PAIR "4 12" [reps]'
set -u
VAR=$1; VALUES=$2; REPS=${3:-2}
python __graft_entry__.py > /dev/null 2>&1 || { echo "build failed"; exit 1; }
for i in $(seq $REPS); do
  for v in $VALUES; do
    env $VAR=$v python bench.py --no-also --no-cpu-baseline --steps 8 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
k = d['kernel_ms_per_step']
print('$VAR=$v', 'images/s', round(d['value'], 2), 'e2e', round(d['e2e']['value'], 2), 'ms', round(d['ms_per_step'], 2), 'head_in', k.get('head_in'), 'step_update', k.get('step_update'), 'ffn', k.get('ffn_fused'), 'qproj', k.get('qproj_fused'), 'out_proj', k.get('out_proj_ln'), 'gather', k.get('msda_gather'), 'MHz', d['clocks']['sm_mhz'])"
  done
done
