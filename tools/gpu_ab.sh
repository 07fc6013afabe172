# A/B of two builds of the library on the same box: bash tools/gpu_ab.sh libA.so libB.so
mkdir -p gpurun_out
for rep in 1 2; do
for lib in "$@"; do
HSC_B200_LIB=$PWD/hierarchical_sparse_coding_b200/$lib timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bench_ab.log 2>&1
python - <<PY
import json
f='gpurun_out/bench_ab.log'
try:
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    print('$lib: value=%.3g k1=%.1f ms k2=%.2f ms e2e=%.3g' % (d['value'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['e2e']['value']))
except Exception as e:
    print(f, 'failed', e); print(open(f).read()[-1500:])
PY
done
done
