mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "n$N rc=$?"
python - <<PY
import json
f='gpurun_out/bench_n$N.log'
try:
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    print('N=$N: value=%.4g ms/step=%.2f k1=%.2f k2=%.2f e2e=%.4g clocks=%s' % (d['value'], d['ms_per_step'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['e2e']['value'], d['clocks']))
except Exception as e:
    print(f, 'failed', e); print(open(f).read()[-2500:])
PY
