mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r1d.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r1d.log').read().strip().splitlines() if l.startswith('{')][-1])
print('value=%.3g k1=%.1f ms k2=%.1f ms (frac %.3f) e2e=%.3g (%d steps) clocks=%s' % (d['value'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['kernels']['k2']['frac'], d['e2e']['value'], d['e2e']['steps'], d['clocks']))
PY
