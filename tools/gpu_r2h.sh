mkdir -p gpurun_out
timeout 300 python tools/prof_c3_levels.py 1 > gpurun_out/prof_c3_levels.log 2>&1; tail -n 8 gpurun_out/prof_c3_levels.log
timeout 300 python tools/prof_c3_levels.py 10 >> gpurun_out/prof_c3_levels.log 2>&1; tail -n 4 gpurun_out/prof_c3_levels.log
for cfg in "ns128_gm8 128 8 1" "ns64_gm1 64 1 1" "ns64_gm8 64 8 1" "ns64_gm8_serial 64 8 0" "ns128_gm1_serial 128 1 0"; do
  set -- $cfg
  HSC_K1_NS=$2 HSC_K1_GRID_MULT=$3 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --pipeline $4 > gpurun_out/bench_r2h_$1.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2h_*.log')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'value=%.4g ms/step=%.2f k1=%.2f k2=%.2f ms e2e=%.4g clocks=%s %s' % (d['value'], d['ms_per_step'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
    except Exception as e:
        print(f, 'no line', e, open(f).read()[-800:])
PY
