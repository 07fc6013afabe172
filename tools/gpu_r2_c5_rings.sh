# config 5 (190 segments, 2 KB rows: at most two CTAs per SM) with deeper stage rings per warp (HSC_K2_RING_KB)
show() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels']
print('$1 ms/step %.2f value %.4g k1 %.2f k2 %.2f clocks %s' % (d['ms_per_step'], d['value'], k['k1_ms'], k['k2_ms'], d['clocks']['sm_mhz']))"; }
for kb in 0 96 128 0 96 128; do
  HSC_K2_RING_KB=$kb timeout 600 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline --no-extra --pipeline 0 2>/dev/null | show "c5 ring_kb=$kb"
done
