# block level of the shared-memory hierarchy (select_smh over slot3): parity, single-sequence latency, config-4 check
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_slot3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_slot3.log
for rep in 1 2; do
  timeout 300 python tools/latency_c1_c3.py 2>&1 | grep -E "c1_cmp|c3 hier"
  timeout 600 python bench.py --workload c2 --steps 2 --warmup 2 --no-cpu-baseline --no-extra --pipeline 0 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('c2 us/atom %.2f' % d['kernels']['us_per_atom_per_signal'])"
done
show() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['kernels']
print('$1 ms/step %.2f value %.4g k1 %.2f k2 %.2f solo k2 %s clocks %s' % (d['ms_per_step'], d['value'], k['k1_ms'], k['k2_ms'], d['roofline'].get('ms_per_launch'), d['clocks']['sm_mhz']))"; }
timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-extra --pipeline 0 2>/dev/null | show "c4 serial"
timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-extra --pipeline 1 2>/dev/null | show "c4 pipe"
timeout 600 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline --no-extra --pipeline 0 2>/dev/null | show "c5"
