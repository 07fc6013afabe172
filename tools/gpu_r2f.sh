# hierarchical encoder on the device (config 3 at north-star tolerances, latency), K1 grid multiplier under the pipeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -rA -k "config3 or hierarchical or dropin or api_errors" > gpurun_out/pytest_gpu_r2f.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|nnz|Error" gpurun_out/pytest_gpu_r2f.log | tail -n 30
timeout 600 python tools/latency_c1_c3.py > gpurun_out/latency_r2f.log 2>&1; tail -n 6 gpurun_out/latency_r2f.log
for gm in 1 4 8; do
  HSC_K1_GRID_MULT=$gm timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --pipeline 1 > gpurun_out/bench_r2f_gm$gm.log 2>&1
done
HSC_K1_GRID_MULT=8 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --pipeline 0 > gpurun_out/bench_r2f_gm8_serial.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2f_*.log')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'value=%.4g ms/step=%.2f k1=%.2f k2=%.2f ms e2e=%.4g clocks=%s %s' % (d['value'], d['ms_per_step'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
    except Exception as e:
        print(f, 'no line', e, open(f).read()[-800:])
PY
