mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "correlate or golden or config1" > gpurun_out/pytest_gpu.log 2>&1; tail -n 2 gpurun_out/pytest_gpu.log
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
HSC_B200_LIB=$PWD/hierarchical_sparse_coding_b200/libhsc_b200_prof.so timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep "K1 tc" | tail -1
