# Round-1 evidence (final kernels): bench line, CPU reference arm, ncu launch list, ncu --set full of K1 (3xFP16) and K2
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r1d.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1d_ref.log 2>&1; echo "ref rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d.csv $CMD > gpurun_out/ncu_launches_r1d.log 2>&1; echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"correlate_tc" -s 1 -c 1 -o gpurun_out/prof_r1d_k1 $CMD > gpurun_out/ncu_full_k1.log 2>&1; echo "k1 rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"pursuit" -s 1 -c 1 -o gpurun_out/prof_r1d_k2 $CMD > gpurun_out/ncu_full_k2.log 2>&1; echo "k2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r1d.log').read().strip().splitlines() if l.startswith('{')][-1])
print('value=%.3g k1=%.1f ms k2=%.1f ms (frac %.3f) e2e=%.3g clocks=%s' % (d['value'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['kernels']['k2']['frac'], d['e2e']['value'], d['clocks']))
PY
