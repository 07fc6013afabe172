# run-to-run determinism of the encoders at the benchmark shapes (a race in the window pipelines / key hierarchy would show up
# as differing events): MP and LoCOMP on config 4 (512 signals) and config 5 (190 segments), 4 runs each, bitwise comparison
timeout 900 python - <<'PY' 2>&1 | tail -12
import sys, hashlib
import numpy as np, torch
sys.path.insert(0, '.')
import bench
import hierarchical_sparse_coding_b200 as hsc
for wl in ('c4', 'c5'):
    w = dict(bench.WORKLOADS[wl])
    D = bench.make_dictionary(w)
    x = bench.make_signals(w, D, seed=1000)
    eng = hsc.Engine(0); eng.set_dictionary(D)
    xd = torch.from_numpy(x).cuda()
    for method in (0, 1):
        opt = eng.make_options(nbNonzeroCoefs=w['atoms'], method=method)
        cap = w['atoms'] * 8 + 256
        digests = []
        for it in range(4):
            evp, evi, evc, states, resid = eng.encode_device(xd, opt, cap)
            torch.cuda.synchronize()
            nb = np.array([s.n_buffered for s in states])
            m = torch.arange(cap, device='cuda')[None, :] < torch.from_numpy(nb).cuda()[:, None]
            h = hashlib.sha256()
            for t in (evp[m], evi[m], evc[m], resid):
                h.update(t.contiguous().cpu().numpy().tobytes())
            digests.append(h.hexdigest()[:16])
        print('%s %s: %d events, digests %s -> %s' % (wl, 'locomp' if method else 'mp', int(nb.sum()), digests, 'IDENTICAL' if len(set(digests)) == 1 else 'DIFFER'), flush=True)
    eng.close()
PY
