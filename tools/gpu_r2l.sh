# early issue of the window bulk loads: parity subset + timing (c4 pipelined / serial, c2, c1/c3 latency)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu_r2l.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_r2l.log
timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/bench_r2l_pipe.log 2>&1
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --pipeline 0 > gpurun_out/bench_r2l_serial.log 2>&1
timeout 600 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu-baseline --pipeline 0 > gpurun_out/bench_r2l_c2.log 2>&1
timeout 300 python tools/latency_c1_c3.py > gpurun_out/latency_r2l.log 2>&1; tail -n 4 gpurun_out/latency_r2l.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2l_*.log')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'value=%.4g ms/step=%.2f k1=%.2f k2=%.2f ms us/atom=%.2f e2e=%.4g e2e+res=%.4g clocks=%s %s' % (d['value'], d['ms_per_step'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['kernels']['us_per_atom_per_signal'], d['e2e']['value'], d['e2e']['with_residual']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
    except Exception as e:
        print(f, 'no line', e, open(f).read()[-1500:])
PY
