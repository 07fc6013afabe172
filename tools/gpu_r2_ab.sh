# Round-2 A/B measurements quoted in DESIGN.md 6 (interleaved runs: consecutive runs heat the board up, single pairs mislead):
#   serial vs streaming pipeline, near-tie re-ranking on / off, K1 co-residency variants (HSC_K1_NS, HSC_K2_SMEM_KB)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_ab_$name.log 2>&1; }
for rep in 1 2; do
  BENCH_ARGS="--pipeline 1" run pipe_$rep HSC_X=0
  BENCH_ARGS="--pipeline 0" run serial_$rep HSC_X=0
  BENCH_ARGS="--pipeline 0 --rerank-tol 0" run serial_norerank_$rep HSC_X=0
  BENCH_ARGS="--pipeline 1 --rerank-tol 0" run pipe_norerank_$rep HSC_X=0
done
BENCH_ARGS="--pipeline 1" run ns64 HSC_K1_NS=64
BENCH_ARGS="--pipeline 1" run ns32_k2cap HSC_K1_NS=32 HSC_K2_SMEM_KB=76
BENCH_ARGS="--pipeline 1" run gm1 HSC_K1_GRID_MULT=1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_ab_*.log')):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'value=%.4g ms/step=%.2f k1=%.2f k2=%.2f ms e2e=%.4g clocks=%s %s' % (
            d['value'], d['ms_per_step'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
    except Exception as e:
        print(f, 'no line', e)
PY
