mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 25 gpurun_out/pytest_gpu.log
timeout 300 python - <<'PY' 2>&1 | tail -12
import numpy as np, torch, sys
sys.path.insert(0, '.')
import bench
import hierarchical_sparse_coding_b200 as hsc
w = dict(bench.WORKLOADS['c4']); w['S'] = 4
D = bench.make_dictionary(w)
x = bench.make_signals(w, D, seed=1000)
eng = hsc.Engine(0); eng.set_dictionary(D)
c = eng.correlate(x).cpu().numpy()
# float64 reference of one signal
xs = np.pad(x[0].astype(np.float64), ((31, 32), (0, 0)))
from numpy.lib.stride_tricks import sliding_window_view
win = sliding_window_view(xs, (64, 4))[:, 0]            # [T, 64, 4]
ref = np.einsum('tlf,klf->tk', win[:8192], D.astype(np.float64))
err = np.abs(c[0][:8192] - ref)
print('max abs err %.3e  rel to max|c| %.3e ; fp32 einsum err %.3e' % (err.max(), err.max() / np.abs(ref).max(),
      np.abs(np.einsum('tlf,klf->tk', win[:8192].astype(np.float32), D) - ref).max()))
PY
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_$name.log 2>&1
  python - <<PY
import json
f='gpurun_out/bench_$name.log'
try:
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    print('$name: value=%.3g k1=%.1f ms k2=%.1f ms (frac %.3f) e2e=%.3g' % (d['value'], d['kernels']['k1_ms'], d['kernels']['k2_ms'], d['kernels']['k2']['frac'], d['e2e']['value']))
except Exception as e:
    print(f, 'failed', e); print(open(f).read()[-1500:])
PY
}
run k1_f16
run k1_tf32 HSC_K1=tf32
HSC_B200_LIB=$PWD/hierarchical_sparse_coding_b200/libhsc_b200_prof.so timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep "K1 tc" | tail -1
