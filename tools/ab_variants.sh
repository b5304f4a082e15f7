#!/bin/bash
# A/B of tuning builds (build/variants/*.so, made by hand with -D overrides) against the in-tree library on the GPU box:
# step time, frame-pairs/s, sad_match time and the remainder (everything else of the step)
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra --unique 32"
for v in "" $(ls build/variants/*.so 2>/dev/null); do
  echo "== ${v:-in-tree}"
  VISO_B200_LIB=${v:+/root/repo/$v} timeout 120 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), round(d['value']), 'sad', round(r['kernel_ms'],3), 'rest', round(d['ms_per_step']-r['kernel_ms'],3), r['queries_left_to_generic_kernel'])"
done
