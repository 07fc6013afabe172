# near-tie re-ranking: what does it cost, fast path vs slow path (HSC_RERANK=1e-12: the watch runs, never fires)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 900 -k "dropin or kmean or full_length" > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_gpu_r2c.log
for wl in c4 c2; do
  for rr in 0 1e-12 4e-6; do
    HSC_RERANK=$rr timeout 600 python bench.py --workload $wl --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r2c_${wl}_$rr.log 2>&1
  done
done
python - <<'PY'
import json
for wl in ('c4','c2'):
    for rr in ('0','1e-12','4e-6'):
        f='bench_r2c_%s_%s' % (wl, rr)
        try:
            d=json.loads([l for l in open('gpurun_out/%s.log' % f).read().strip().splitlines() if l.startswith('{')][-1])
            print(f, 'value=%.4g ms/step=%.2f k1=%.2f k2=%.2f ms reranked=%s us/atom=%.2f clocks=%s' % (d['value'], d['ms_per_step'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['run']['reranked_selections_per_step'], d['kernels']['us_per_atom_per_signal'], d['clocks']['sm_mhz']))
        except Exception as e:
            print(f, 'no line', e)
PY
