# edge-path window variant (parity tests that crowd the signal ends), then the pipeline with a high-priority K1 stream
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -k "edge or random_cases or shape_sweep or degenerate or golden_mp or variants or full_length" > gpurun_out/pytest_gpu_r2i.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu_r2i.log
for cfg in "p0_ns128_gm8 0 128 8 2" "p1_ns128_gm8 1 128 8 2" "p1_ns64_gm8 1 64 8 2" "p1_ns64_gm1 1 64 1 2" "p1_ns64_gm8_s3 1 64 8 3" "p1_ns96_gm8 1 96 8 2"; do
  set -- $cfg
  HSC_K1_NS=$3 HSC_K1_GRID_MULT=$4 timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --pipeline 1 --k1-priority $2 --slots $5 > gpurun_out/bench_r2i_$1.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2i_*.log')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'value=%.4g ms/step=%.2f k1=%.2f k2=%.2f ms e2e=%.4g clocks=%s %s' % (d['value'], d['ms_per_step'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
    except Exception as e:
        print(f, 'no line', e, open(f).read()[-800:])
PY
