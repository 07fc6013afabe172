# eight GPUs: bench under torchrun (weak scaling: 512 signals per GPU), host topology for the e2e analysis
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/n8_topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|^CPU\(s\)|Model name" > gpurun_out/n8_lscpu.txt 2>&1
for N in 8 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 6 --warmup 3 --no-extra > gpurun_out/bench_r2_n$N.log 2>&1; echo "bench n$N rc=$?"
done
python - <<'PY'
import json
for N in (8, 2):
    try:
        d=json.loads([l for l in open('gpurun_out/bench_r2_n%d.log' % N).read().strip().splitlines() if l.startswith('{')][-1])
        print('N=%d value=%.4g ms/step=%.2f e2e=%.4g e2e+res=%.4g clocks=%s' % (N, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['with_residual']['value'], d['clocks']))
    except Exception as e:
        print('no line', e, open('gpurun_out/bench_r2_n%d.log' % N).read()[-2000:])
PY
cat gpurun_out/n8_lscpu.txt; head -n 14 gpurun_out/n8_topo.txt
