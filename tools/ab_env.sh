#!/bin/bash
# A/B of run-time tuning knobs (environment variables) on the GPU box: usage  tools/ab_env.sh VAR=a VAR=b "VAR=c OTHER=d" ...
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra --unique 32"
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 160 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), round(d['value']), 'sad', round(r['kernel_ms'],3), 'rest', round(d['ms_per_step']-r['kernel_ms'],3), 'ok pairs', d['poses']['ok_frame_pairs'], 'inliers', round(d['poses']['inliers_mean'],3))"
done
