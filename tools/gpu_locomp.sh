mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tail -8
import numpy as np, torch, sys, time
sys.path.insert(0, '.')
import bench
import hierarchical_sparse_coding_b200 as hsc
for wl, S in (('c4', 512), ('c5', 190)):
    w = dict(bench.WORKLOADS[wl]); w['S'] = S
    D = bench.make_dictionary(w)
    x = bench.make_signals(w, D, seed=1000)
    eng = hsc.Engine(0); eng.set_dictionary(D)
    xd = torch.from_numpy(x).cuda()
    for method in (0, 1):
        opt = eng.make_options(nbNonzeroCoefs=w['atoms'], method=method)
        cap = w['atoms'] * 8 + 256
        for it in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            evp, evi, evc, states, resid = eng.encode_device(xd, opt, cap)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ne = sum(s.n_events for s in states); nnz = sum(s.nnz for s in states)
        stops = {}
        for s in states: stops[s.status] = stops.get(s.status, 0) + 1
        e_sig = sum(s.energy_signal for s in states); e_res = sum(s.energy_residual for s in states)
        print('%s method=%s: %.1f ms  events %.1f/signal nnz %.1f/signal  %.3g atoms/s  SNR %.2f dB stops %s' % (
            wl, 'locomp' if method else 'mp', 1e3 * dt, ne / S, nnz / S, nnz / dt, 10 * np.log10(e_sig / e_res), stops))
    eng.close()
PY
