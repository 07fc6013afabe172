mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -rA -k "full_length_locomp" > gpurun_out/pytest_gpu_r2q.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|reference events|Error|assert" gpurun_out/pytest_gpu_r2q.log | tail -n 12
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2q.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_r2q.log').read().strip().splitlines() if l.startswith('{')][-1])
    print('value=%.4g ms/step=%.2f k2 pipe %.2f solo k1 %.2f k2 %.2f (frac %.3f) e2e %.4g / %.4g cpu %.1f' % (d['value'], d['ms_per_step'], d['kernels']['k2_ms'], d['kernels']['k1']['solo_ms_per_launch'], d['kernels']['k2']['solo_ms_per_launch'], d['kernels']['k2']['solo_frac'], d['e2e']['value'], d['e2e']['with_residual']['value'], d['cpu_baseline']['value']))
    print(json.dumps(d.get('extra'), indent=1)[:3000])
except Exception as e:
    print('no line', e, open('gpurun_out/bench_r2q.log').read()[-3000:])
PY
