mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -n 30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c4_b.log 2>&1
python - <<'PY'
import json
f='gpurun_out/bench_c4_b.log'
try:
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    print(f, 'value=%.3g k1=%.1f ms (%.1f TF/s useful, frac %.3f) k2=%.1f ms k2frac=%.3f e2e=%.3g' % (d['value'], d['kernels']['k1_ms'], d['kernels']['k1']['achieved'], d['kernels']['k1']['frac'], d['kernels']['k2_ms'], d['kernels']['k2']['frac'], d['e2e']['value']))
except Exception as e:
    print(f, 'failed', e); print(open(f).read()[-1500:])
PY
