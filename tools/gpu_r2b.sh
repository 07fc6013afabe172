# Round 2, second GPU pass: parity suite + drop-in tests, bench (c4, c2) with the near-tie re-ranking on / off
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -rA > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|near-tie|step-identical|common prefix" gpurun_out/pytest_gpu_r2b.log | tail -n 60
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2b.log 2>&1; echo "bench rc=$?"
HSC_RERANK=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2b_norerank.log 2>&1; echo "bench(no rerank) rc=$?"
timeout 600 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r2b_c2.log 2>&1; echo "bench c2 rc=$?"
HSC_RERANK=0 timeout 600 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r2b_c2_norerank.log 2>&1
python - <<'PY'
import json
for f in ('bench_r2b', 'bench_r2b_norerank', 'bench_r2b_c2', 'bench_r2b_c2_norerank'):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.log' % f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value=%.4g ms/step=%.2f k1=%.2f ms k2=%.2f ms (frac %.3f) e2e=%.4g' % (d['value'], d['ms_per_step'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['kernels']['k2']['frac'], d['e2e']['value']))
    except Exception as e:
        print(f, 'no line', e)
PY
