# wide rows (32 lanes per row) through the lean window loops (HSC_K2_ROW32 = 0 general loop, 1 gram_update_row32): parity, then
# interleaved A/B on configs 4 and 5
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_row32.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_row32.log
show() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels']
print('$1 ms/step %.2f value %.4g k1 %.2f k2 %.2f solo k2 %s clocks %s' % (d['ms_per_step'], d['value'], k['k1_ms'], k['k2_ms'], d['roofline'].get('ms_per_launch'), d['clocks']['sm_mhz']))"; }
for r in 0 1 0 1; do
  echo "== row32 $r"
  HSC_K2_ROW32=$r timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-extra --pipeline 0 2>/dev/null | show "c4 serial"
  HSC_K2_ROW32=$r timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-extra --pipeline 1 2>/dev/null | show "c4 pipe"
  HSC_K2_ROW32=$r timeout 600 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline --no-extra --pipeline 0 2>/dev/null | show "c5"
done
