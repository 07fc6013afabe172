# where the K-SVD loop's coefficient stage spends its time (config 5 segments): set_dictionary / encode / code accumulation
mkdir -p gpurun_out
timeout 900 python - <<'PY' 2>&1 | tail -40
import sys, time, math
import numpy as np, torch
sys.path.insert(0, '.')
import bench
import hierarchical_sparse_coding_b200 as hsc
from hierarchical_sparse_coding_b200.modeling import get_engine
w = dict(bench.WORKLOADS['c5'])
D = bench.make_dictionary(w)
x = bench.make_signals(w, D, seed=1000)
eng = get_engine(0)
xd = torch.from_numpy(x).cuda(); resid = torch.empty_like(xd)
rs = np.random.RandomState(7)
S, T = x.shape[0], x.shape[1]
for it in range(10):
    Dn = D.astype(np.float64) + 0.05 * rs.randn(*D.shape); Dn /= np.sqrt(np.sum(Dn * Dn, axis=(1, 2), keepdims=True))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.set_dictionary(Dn, dtype=np.float32)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    opt = eng.make_options(w['atoms'], None, None, 1, 1e-16)
    cap = eng.default_capacity(opt, T)
    evp, evi, evc, st_arr, _ = eng.encode_device(xd, opt, cap, resid=resid)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    states = [st_arr[i] for i in range(S)]
    nb = torch.tensor([st.n_buffered for st in states], dtype=torch.int64, device='cuda')
    mask = torch.arange(evp.shape[1], device='cuda')[None, :] < nb[:, None]
    sig = torch.arange(S, device='cuda')[:, None].expand(S, evp.shape[1])[mask]
    pos, idx, coef = evp[mask], evi[mask], evc[mask]
    torch.cuda.synchronize(); t3 = time.perf_counter()
    sg, p, ix, c, col_ptr = eng.accumulate_code(sig, pos, idx, coef, S, T, D.shape[0], 1e-16)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    Dk, c_new, alpha = eng.ksvd_update(Dn, sg, p, ix, c, col_ptr, S, T)
    torch.cuda.synchronize(); t5 = time.perf_counter()
    print('iter %d: set_dictionary %.1f ms, encode %.1f ms, select events %.1f ms, accumulate %.1f ms, update %.1f ms; torch reserved %.1f GB' % (
        it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4), torch.cuda.memory_reserved() / 1e9), flush=True)
PY
