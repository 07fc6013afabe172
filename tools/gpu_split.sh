# vectorised fp16 split pre-pass: parity suite, bench line, and which split kernel ran (ncu, 3 launches)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_split.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_split.log').read().strip().splitlines() if l.startswith('{')][-1])
print('value=%.3g k1=%.2f ms k2=%.1f ms e2e=%.3g' % (d['value'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['e2e']['value']))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:split -c 1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep -E "split_signal|gpu__time" | head -6
