mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest (v6 default) rc=$?"
tail -n 3 gpurun_out/pytest_gpu.log
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_$name.log 2>&1
  python - <<PY
import json
f='gpurun_out/bench_$name.log'
try:
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    print('$name: value=%.3g k1=%.1f ms k2=%.1f ms (frac %.3f) e2e=%.3g' % (d['value'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['kernels']['k2']['frac'], d['e2e']['value']))
except Exception as e:
    print(f, 'failed', e); print(open(f).read()[-1500:])
PY
}
run v6 HSC_PURSUIT_VARIANT=6
run v4 HSC_PURSUIT_VARIANT=4
run v6_s2 HSC_PURSUIT_VARIANT=6 HSC_K2_TMA_STAGES=2
run v6_again HSC_PURSUIT_VARIANT=6
HSC_PURSUIT_VARIANT=6 bash tools/gpu_phases.sh; grep "hsc " gpurun_out/prof_dump.log | tail -4
