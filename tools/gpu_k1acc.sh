mkdir -p gpurun_out
for mode in f16 tf32 simt; do
HSC_K1=$mode timeout 300 python - <<'PY' 2>&1 | tail -3
import numpy as np, torch, sys, os
sys.path.insert(0, '.')
import bench
import hierarchical_sparse_coding_b200 as hsc
from numpy.lib.stride_tricks import sliding_window_view
for wl, S in (('c4', 2), ('c5', 2), ('c2', 1)):
    w = dict(bench.WORKLOADS[wl]); w['S'] = S
    if wl == 'c2': w['T'] = 65536; w['atoms'] = 600
    D = bench.make_dictionary(w)
    x = bench.make_signals(w, D, seed=1000)
    eng = hsc.Engine(0); eng.set_dictionary(D)
    c = eng.correlate(x).cpu().numpy()
    L, F = w['L'], w['F']; off = L // 2 - 1
    xs = np.pad(x[0].astype(np.float64), ((off, L - 1 - off), (0, 0)))
    win = sliding_window_view(xs, (L, F))[:, 0]
    n = 4096
    ref = np.einsum('tlf,klf->tk', win[:n], D.astype(np.float64))
    f32 = np.einsum('tlf,klf->tk', win[:n].astype(np.float32), D)
    err = np.abs(c[0][:n] - ref)
    print(os.environ.get('HSC_K1'), wl, 'max abs err %.3e  rel to max|c| %.3e  rms %.3e ; numpy fp32 max err %.3e rms %.3e' % (
        err.max(), err.max() / np.abs(ref).max(), np.sqrt(np.mean(err**2)), np.abs(f32 - ref).max(), np.sqrt(np.mean((f32-ref)**2))))
PY
done
