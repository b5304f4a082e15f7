B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra --unique 32"
for v in "" $(ls build/variants/*.so); do
  echo "== $v"
  VISO_B200_LIB=${v:+/root/repo/$v} timeout 120 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), round(d['value']), round(r['kernel_ms'],3), r['sad_pairs_per_step'], r['queries_left_to_generic_kernel'])"
done
