"""Where the GPU time of the hierarchical encode (config 3) goes: CUDA-event times of K1 (begin) and K2 (run) per level."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hierarchical_sparse_coding_b200 as hsc
from hierarchical_sparse_coding_b200.engine import engine_for_dictionary, engine_dtype
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'c3_complex.npz'))
nl = int(z['nb_levels'])
raw = [z['raw_l%d' % l] for l in range(nl)]
cns = z['counts_no_singletons']
x = z['x']
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for rep in range(3):
    xd = None
    line = []
    for level in range(nl):
        D = raw[level]
        w = np.ones((D.shape[0],), D.dtype); w[:D.shape[0] - cns[level]] = 0.95
        if xd is None:
            dt = engine_dtype(x[:, None], D)
            eng = engine_for_dictionary(D, w, dt)
            xd = torch.from_numpy(np.ascontiguousarray(x[None, :, None], dtype=dt)).cuda()
        else:
            eng = engine_for_dictionary(D, w, np.float64)
        opt = eng.make_options(None, None, 10.0, nb, 1e-16, use_weights=True)
        cap = eng.default_capacity(opt, x.shape[0])
        resid = torch.empty_like(xd)
        evp = torch.empty((1, cap), dtype=torch.int32, device='cuda'); evi = torch.empty_like(evp)
        evc = torch.empty((1, cap), dtype=eng.torch_dtype, device='cuda')
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda.synchronize()
        e[0].record(); eng.begin_only(xd, opt, resid); e[1].record()
        st = eng.run_only(evp, evi, evc, cap, sync_states=False); e[2].record()
        xd2 = eng.events_to_dense(evp, evi, evc, 1e-16); e[3].record()
        torch.cuda.synchronize()
        import ctypes
        from hierarchical_sparse_coding_b200 import _native as N
        states = (N.SignalState * 1)()
        N.check(eng.lib, eng.handle, eng.lib.hsc_b200_mp_states(eng.handle, states, eng._stream_ptr()))
        line.append('L%d[%s T=%d F=%d K=%d L=%d]: K1 %.2f ms, K2 %.2f ms (%d atoms, %.1f us/atom), dense %.2f ms' % (
            level, np.dtype(eng.dtype).name, xd.shape[1], eng.F, eng.K, eng.L, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]),
            states[0].n_events, 1e3 * e[1].elapsed_time(e[2]) / max(states[0].n_events, 1), e[2].elapsed_time(e[3])))
        xd = xd2
    print('nbBlocks=%d rep %d\n  ' % (nb, rep) + '\n  '.join(line))
