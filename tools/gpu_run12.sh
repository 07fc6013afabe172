mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; tail -n 5 gpurun_out/pytest_gpu.log
for ch in 1 2 4 8; do
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --chunks $ch > gpurun_out/bench_c4_ch$ch.log 2>&1
python - <<PY
import json
f='gpurun_out/bench_c4_ch$ch.log'
try:
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    print('chunks $ch: value=%.3g k1=%.1f ms k2=%.1f ms e2e=%.3g' % (d['value'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['e2e']['value']))
except Exception as e:
    print(f, 'failed', e); print(open(f).read()[-1500:])
PY
done
