# streaming pipeline vs serial steps; near-tie re-ranking on/off interleaved (clock drift between runs confounds single A/B pairs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -k "pipelined or c_abi or full_length or batch_equals" > gpurun_out/pytest_gpu_r2e.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_gpu_r2e.log
for rep in 1 2; do
  for cfg in "p1_rr 1 -1" "p0_rr 0 -1" "p0_norr 0 0" "p1_norr 1 0"; do
    set -- $cfg
    timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --pipeline $2 --rerank-tol $3 > gpurun_out/bench_r2e_$1_$rep.log 2>&1
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2e_*.log')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'value=%.4g ms/step=%.2f k1=%.2f k2=%.2f ms reranked=%s e2e=%.4g e2e+res=%.4g clocks=%s %s' % (d['value'], d['ms_per_step'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['run']['reranked_selections_per_step'], d['e2e']['value'], d['e2e']['with_residual']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
    except Exception as e:
        print(f, 'no line', e, open(f).read()[-800:])
PY
