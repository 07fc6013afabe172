mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
for v in 0 1 2 3; do
  HSC_PURSUIT_VARIANT=$v timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_v$v.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_v$v.log').read().strip().splitlines()[-1])
    print('variant $v: value=%.3g atoms/s k1=%.1f ms k2=%.1f ms k2frac=%.3f e2e=%.3g' % (d['value'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['kernels']['k2']['frac'], d['e2e']['value']))
except Exception as e:
    print('variant $v failed', e)
PY
done
