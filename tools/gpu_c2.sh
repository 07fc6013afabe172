mkdir -p gpurun_out
for smh in 0 1; do
HSC_K2_SMH=$smh timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_c2.log').read().strip().splitlines() if l.startswith('{')][-1])
print('c2 smh=$smh value=%.3g k1=%.2f ms k2=%.2f ms us/atom=%.2f e2e=%.3g' % (d['value'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['kernels']['us_per_atom_per_signal'], d['e2e']['value']))
PY
done
timeout 600 python -m pytest tests -m gpu -q --timeout 500 -x -k "config2 or config1 or golden_mp or variants" 2>&1 | tail -2
