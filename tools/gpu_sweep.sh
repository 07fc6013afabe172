# K2 time against the number of signals per GPU (1..4 CTAs per SM): chain-bound vs memory-bound
mkdir -p gpurun_out
for S in 74 148 296 444 512 592; do
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --signals $S > gpurun_out/bench_sweep.log 2>&1
python - <<PY
import json
f='gpurun_out/bench_sweep.log'
try:
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    print('S=$S: value=%.3g k1=%.2f ms k2=%.2f ms  atoms/signal=%.1f  k2 GB/s(alg)=%.0f e2e=%.3g' % (d['value'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['config']['selections_per_signal'], d['kernels']['k2']['achieved'], d['e2e']['value']))
except Exception as e:
    print(f, 'failed', e); print(open(f).read()[-800:])
PY
done
