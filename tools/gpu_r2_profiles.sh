# Round-2 evidence set: bench line (N=1, with the CPU baseline), CPU arm through the real reference, ncu launch list of the
# bench command, ncu --set full of K1 and K2 on config 4 and of K2 on the config-2 / config-5 shapes.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rs > gpurun_out/pytest_gpu_r2.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_r2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python bench.py > gpurun_out/bench_r2.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2_ref.log 2>&1; echo "ref rc=$?"
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/bench_r2_c2.log 2>&1; echo "c2 rc=$?"
timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 --ksvd-iters 8 > gpurun_out/bench_r2_c5.log 2>&1; echo "c5 rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_launches_r2.log 2>&1; echo "launch list rc=$?"
CMD1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --pipeline 0"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"correlate_tc" -s 3 -c 1 -o gpurun_out/prof_r2_k1 $CMD1 > gpurun_out/ncu_full_k1.log 2>&1; echo "k1 rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"pursuit_kernel" -s 3 -c 1 -o gpurun_out/prof_r2_k2 $CMD1 > gpurun_out/ncu_full_k2.log 2>&1; echo "k2 rc=$?"
timeout 1500 ncu --set full --clock-control none -k regex:"pursuit_kernel" -s 3 -c 1 -o gpurun_out/prof_r2_k2_c2 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --pipeline 0 > gpurun_out/ncu_full_k2_c2.log 2>&1; echo "k2 c2 rc=$?"
timeout 1500 ncu --set full --clock-control none -k regex:"pursuit_kernel" -s 3 -c 1 -o gpurun_out/prof_r2_k2_c5 python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline --pipeline 0 > gpurun_out/ncu_full_k2_c5.log 2>&1; echo "k2 c5 rc=$?"
timeout 1500 ncu --set full --clock-control none -k regex:"correlate_tc" -s 3 -c 1 -o gpurun_out/prof_r2_k1_c5 python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline --pipeline 0 > gpurun_out/ncu_full_k1_c5.log 2>&1; echo "k1 c5 rc=$?"
timeout 300 python tools/locomp_c4.py > gpurun_out/locomp_c4.log 2>&1; echo "locomp plain rc=$?"; tail -1 gpurun_out/locomp_c4.log
timeout 900 ncu --set full --clock-control none -k regex:"locomp" -s 1 -c 1 -f -o gpurun_out/prof_r2_locomp python tools/locomp_c4.py > gpurun_out/ncu_full_locomp.log 2>&1; echo "locomp ncu rc=$?"
bash tools/gpu_locomp.sh > gpurun_out/locomp_mp_c4_c5.log 2>&1; cat gpurun_out/locomp_mp_c4_c5.log
# summaries on the box (the merge back is capped at 64 MiB: only the two config-4 reports travel)
for r in prof_r2_k1 prof_r2_k2 prof_r2_k2_c2 prof_r2_k2_c5 prof_r2_k1_c5 prof_r2_locomp; do
  python tools/ncu_summary.py gpurun_out/$r.ncu-rep > gpurun_out/$r.json 2>/dev/null
done
ncu -i gpurun_out/prof_r2_k2.ncu-rep --page source --csv > gpurun_out/prof_r2_k2_source.csv 2>/dev/null
python - <<'PY2'
# the 40 hottest source lines of K2 (instructions executed), for the README of profiles/
import csv
try:
    rows = list(csv.reader(open('gpurun_out/prof_r2_k2_source.csv', errors='replace')))
    hdr = rows[0]
    col = [i for i, h in enumerate(hdr) if 'Instructions Executed' in h][:1]
    src = [i for i, h in enumerate(hdr) if h.strip() in ('Source',)][:1]
    if col and src:
        def num(v):
            try: return float(v.replace(',', ''))
            except Exception: return 0.0
        top = sorted(rows[1:], key=lambda r: -num(r[col[0]]))[:40]
        with open('gpurun_out/prof_r2_k2_hot_lines.txt', 'w') as f:
            for r in top:
                f.write('%14s  %s\n' % (r[col[0]], r[src[0]][:160]))
except Exception as e:
    print('source page not summarised:', e)
PY2
rm -f gpurun_out/prof_r2_k2_c2.ncu-rep gpurun_out/prof_r2_k2_c5.ncu-rep gpurun_out/prof_r2_k1_c5.ncu-rep gpurun_out/prof_r2_locomp.ncu-rep gpurun_out/prof_r2_k2_source.csv
ls -la gpurun_out/ | head -40; du -sh gpurun_out
python - <<'PY'
import json
for f in ('bench_r2', 'bench_r2_ref', 'bench_r2_c2', 'bench_r2_c5'):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.log' % f).read().strip().splitlines() if l.startswith('{')][-1])
        if 'kernels' in d:
            print(f, 'value=%.4g ms/step=%.2f k1=%.2f k2=%.2f ms e2e=%.4g clocks=%s' % (d['value'], d['ms_per_step'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['e2e']['value'], d['clocks']), (d.get('cpu_baseline') or {}).get('value'), (d.get('extra') or {}))
        else:
            print(f, d['value'], d['ms_per_step'])
    except Exception as e:
        print(f, 'no line', e, open('gpurun_out/%s.log' % f).read()[-1500:])
PY
