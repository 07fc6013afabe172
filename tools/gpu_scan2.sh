mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; tail -n 2 gpurun_out/pytest_gpu.log
for pf in 0 1; do for v in 1 4; do for S in 148 512; do
  HSC_PREFETCH=$pf HSC_PURSUIT_VARIANT=$v timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --signals $S > gpurun_out/scan.log 2>&1
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/scan.log').read().strip().splitlines() if l.startswith('{')][-1])
    print('prefetch $pf variant $v S=$S: k1=%.1f ms k2=%.1f ms k2 atoms/us=%.2f k2frac=%.3f' % (d['kernels']['k1_ms'], d['kernels']['k2_ms'], $S*d['config']['selections_per_signal']/d['kernels']['k2_ms']/1e3, d['kernels']['k2']['frac']))
except Exception as e:
    print('variant $v S=$S failed', e)
PY
done; done; done
