# two GPUs: the multi-GPU correctness tests (skipped by the single-GPU round-end run) and the bench under torchrun
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -rA -k "distributed_ksvd or sharded_encode" > gpurun_out/pytest_gpu_r2_n2.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|PASSED|FAILED|single process|single GPU" gpurun_out/pytest_gpu_r2_n2.log | tail -n 12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_r2_n2.log 2>&1; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_r2_n2.log').read().strip().splitlines() if l.startswith('{')][-1])
    print('N=2 value=%.4g ms/step=%.2f e2e=%.4g e2e+res=%.4g clocks=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['with_residual']['value'], d['clocks']))
except Exception as e:
    print('no line', e, open('gpurun_out/bench_r2_n2.log').read()[-2000:])
PY
