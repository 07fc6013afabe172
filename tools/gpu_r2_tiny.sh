# narrow maps on the register window path under the shared-memory hierarchy (HSC_K2_TINYROW): parity + latency A/B
mkdir -p gpurun_out
echo skip tests
for tr in 0 1 2 0 1 2; do
  echo "== tiny row $tr"
  HSC_K2_TINYROW=$tr timeout 300 python tools/latency_c1_c3.py 2>&1 | grep -E "c1_cmp|c3 hier"
  HSC_K2_TINYROW=$tr timeout 600 python bench.py --workload c2 --steps 2 --warmup 2 --no-cpu-baseline --pipeline 0 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('c2 us/atom %.2f' % d['kernels']['us_per_atom_per_signal'])"
done
