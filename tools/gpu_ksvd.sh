mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "ksvd" 2>&1 | tail -3
timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline --ksvd-iters 7 > gpurun_out/bench_c5.log 2>&1; echo "c5 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_c5.log').read().strip().splitlines() if l.startswith('{')][-1])
print('c5 value=%.3g' % d['value'], [(round(h['encode_s'],3), round(h['update_s'],3)) for h in d['extra']['ksvd']['history']])
PY
