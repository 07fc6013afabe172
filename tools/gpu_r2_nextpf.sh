# next-pick prefetch from the watch warp (HSC_K2_NEXT_PREFETCH), with the group of the runner-up found through the block scores
show() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels']
print('$1 ms/step %.2f value %.4g k1 %.2f k2 %.2f clocks %s' % (d['ms_per_step'], d['value'], k['k1_ms'], k['k2_ms'], d['clocks']['sm_mhz']))"; }
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "full_length_reference or golden_mp" 2>&1 | tail -2
for rep in 1 2; do
  for pf in 0 1; do
    HSC_K2_NEXT_PREFETCH=$pf timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-extra --pipeline 0 2>/dev/null | show "prefetch=$pf c4 serial"
    HSC_K2_NEXT_PREFETCH=$pf timeout 600 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline --no-extra --pipeline 0 2>/dev/null | show "prefetch=$pf c5"
  done
done
for pf in 0 1; do HSC_K2_NEXT_PREFETCH=$pf timeout 600 python bench.py --workload c2 --steps 2 --warmup 2 --no-cpu-baseline --no-extra --pipeline 0 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('prefetch=$pf c2 us/atom %.2f' % d['kernels']['us_per_atom_per_signal'])"; done
