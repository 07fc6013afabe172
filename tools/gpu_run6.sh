mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
for v in 1 0 2 3; do
  HSC_PURSUIT_VARIANT=$v timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_v$v.log 2>&1
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_v$v.log').read().strip().splitlines() if l.startswith('{')][-1])
    print('variant $v: value=%.3g atoms/s k1=%.1f ms k2=%.1f ms k2frac=%.3f e2e=%.3g' % (d['value'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['kernels']['k2']['frac'], d['e2e']['value']))
except Exception as e:
    print('variant $v failed', e)
PY
done
CMD="python bench.py --steps 1 --warmup 1 --signals 148 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain148.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"correlate_tc" -s 1 -c 1 -o gpurun_out/prof_k1tc $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
