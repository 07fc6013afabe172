mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu_r2p.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_r2p.log
for ei in 0 1 0 1; do
  echo "== early issue $ei"
  HSC_K2_EARLY_ISSUE=$ei timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --pipeline 0 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('c4 serial ms/step %.2f k2 %.2f clocks %s' % (d['ms_per_step'], d['kernels']['k2_ms'], d['clocks']['sm_mhz']))"
  HSC_K2_EARLY_ISSUE=$ei timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --pipeline 1 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('c4 pipe ms/step %.2f k2 %.2f solo %.2f clocks %s e2e %.4g' % (d['ms_per_step'], d['kernels']['k2_ms'], d['kernels']['k2']['solo_ms_per_launch'], d['clocks']['sm_mhz'], d['e2e']['value']))"
  HSC_K2_EARLY_ISSUE=$ei timeout 600 python bench.py --workload c2 --steps 2 --warmup 2 --no-cpu-baseline --pipeline 0 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('c2 us/atom %.2f' % d['kernels']['us_per_atom_per_signal'])"
  HSC_K2_EARLY_ISSUE=$ei timeout 300 python tools/latency_c1_c3.py 2>&1 | grep -E "c1_cmp|c3 hier"
done
