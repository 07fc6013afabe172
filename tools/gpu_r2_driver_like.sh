# what the driver runs at round end, with its arguments: smoke, pytest -m gpu, bench (both arms) at --steps 20 --warmup 3
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_r2.log 2>&1; tail -n 4 gpurun_out/smoke_r2.log
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_r2_final.log 2>&1; tail -n 6 gpurun_out/pytest_gpu_r2_final.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 ) > gpurun_out/bench_r2_driver_ref.log 2>&1; tail -n 4 gpurun_out/bench_r2_driver_ref.log | cut -c1-400
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 ) > gpurun_out/bench_r2_driver.log 2>&1; tail -n 4 gpurun_out/bench_r2_driver.log | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_driver.log').read().strip().splitlines() if l.startswith('{')][-1])
print('value=%.4g ms/step=%.2f k2 pipe %.2f solo k1 %.2f k2 %.2f (frac %.3f) e2e %.4g / %.4g cpu %.1f launches %d' % (d['value'], d['ms_per_step'], d['kernels']['k2_ms'], d['kernels']['k1_solo_ms'], d['kernels']['k2_solo_ms'], d['kernels']['k2']['frac'], d['e2e']['value'], d['e2e']['with_residual']['value'], d['cpu_baseline']['value'], d['gpu_launches']))
print(json.dumps(d['extra']['other_configs']['c5'].get('ksvd_loop')), d['extra']['other_configs']['c2'].get('us_per_selection_per_signal'))
print('roofline', {k: d['roofline'][k] for k in ('kernel','achieved','peak','frac','ms_per_launch','traffic') if k in d['roofline']})
r=json.loads([l for l in open('gpurun_out/bench_r2_driver_ref.log').read().strip().splitlines() if l.startswith('{')][-1])
print('ref', r['value'], r['ms_per_step'], r['config']==d['config'], r['cpu_baseline']['cores'], r['cpu_baseline']['nb_nonzero_coefs_per_step'])
PY
