mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_now.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_now.log').read().strip().splitlines() if l.startswith('{')][-1])
print('value=%.3g k1=%.1f ms k2=%.1f ms (frac %.3f) e2e=%.3g' % (d['value'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['kernels']['k2']['frac'], d['e2e']['value']))
PY
