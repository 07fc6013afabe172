mkdir -p gpurun_out
for cfg in "base 128 8 0 2" "ns64 64 8 0 2" "k2s76_ns32 32 8 76 2" "k2s57_ns64 64 8 57 2" "base_serial 128 1 0 2"; do
  set -- $cfg
  P=1; if [ "$1" = "base_serial" ]; then P=0; fi
  HSC_K1_NS=$2 HSC_K1_GRID_MULT=$3 HSC_K2_SMEM_KB=$4 timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --pipeline $P --slots $5 > gpurun_out/bench_r2k_$1.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2k_*.log')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'value=%.4g ms/step=%.2f k1=%.2f k2=%.2f ms e2e=%.4g clocks=%s %s' % (d['value'], d['ms_per_step'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
    except Exception as e:
        print(f, 'no line', e, open(f).read()[-800:])
PY
