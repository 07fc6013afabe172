mkdir -p gpurun_out
for st in 3 6 3 6; do
timeout 300 python bench.py --steps $st --warmup 2 --no-cpu-baseline > gpurun_out/bench_e2e.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_e2e.log').read().strip().splitlines() if l.startswith('{')][-1])
print('steps=$st value=%.3g e2e=%.3g (e2e steps %d)' % (d['value'], d['e2e']['value'], d['e2e']['steps']))
PY
done
timeout 300 python - <<'PY'
import numpy as np, torch, sys, time
sys.path.insert(0, '.')
import bench
import hierarchical_sparse_coding_b200 as hsc
w = dict(bench.WORKLOADS['c4'])
D = bench.make_dictionary(w)
x = bench.make_signals(w, D, seed=1000)
eng = hsc.Engine(0); eng.set_dictionary(D)
opt = eng.make_options(nbNonzeroCoefs=w['atoms'])
cap = w['atoms'] * 4 + 64
x_pin = torch.from_numpy(x).pin_memory()
outs = [torch.empty_like(x_pin).pin_memory() for _ in range(2)]
for rep in range(2):
    t0 = time.perf_counter(); ts = []
    for r in eng.encode_host_pipelined((x_pin for _ in range(8)), opt, cap, n_chunks=8, residual_outs=[outs[i & 1] for i in range(8)]):
        ts.append(time.perf_counter() - t0)
    print('yield times (ms):', ' '.join('%.1f' % (1e3 * t) for t in ts))
# raw PCIe rates on this box
a = torch.empty_like(x_pin, device='cuda')
for name, fn in (('h2d', lambda: a.copy_(x_pin, non_blocking=True)), ('d2h', lambda: outs[0].copy_(a, non_blocking=True))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(name, '%.1f GB/s' % (x_pin.numel() * 4 / dt / 1e9))
PY
