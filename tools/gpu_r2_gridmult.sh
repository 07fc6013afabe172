show() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels']
print('$1 ms/step %.2f value %.4g k1 %.2f k2 %.2f clocks %s' % (d['ms_per_step'], d['value'], k['k1_ms'], k['k2_ms'], d['clocks']['sm_mhz']))"; }
for rep in 1 2; do
for gm in 8 4 16 2; do
  HSC_K1_GRID_MULT=$gm timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-extra --pipeline 1 2>/dev/null | show "grid_mult=$gm pipe"
done
done
