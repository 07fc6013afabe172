mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c4_now.log 2>&1; echo "c4 rc=$?"
timeout 900 python bench.py --workload c5 --steps 3 --warmup 2 --no-cpu-baseline --ksvd-iters 3 > gpurun_out/bench_c5.log 2>&1; echo "c5 rc=$?"
timeout 900 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1; echo "c2 rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/bench_c4_now.log','gpurun_out/bench_c5.log','gpurun_out/bench_c2.log']:
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value=%.3g k1=%.2f ms k2=%.2f ms (frac %.3f) e2e=%.3g' % (d['value'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['kernels']['k2']['frac'], d['e2e']['value']), d.get('extra'))
    except Exception as e:
        print(f, 'failed', e); print(open(f).read()[-2500:])
PY
