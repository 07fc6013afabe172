import sys, os
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
from helpers import load_npz, case_kwargs, TraceComparison, snr_db
import hierarchical_sparse_coding_b200 as hsc
z = load_npz('mp_cases.npz')
names = [str(n) for n in z['names'] if str(z[str(n) + '_method']) == 'locomp']
for name in names:
    kw = case_kwargs(z, name)
    x, D = z[name + '_x'], z[name + '_D']
    cmp = hsc.LoCOMP()
    try:
        coef, res = cmp.computeCoefficients(x, D, **kw)
    except Exception as e:
        print(name, 'EXC', e); continue
    r = cmp.last_result
    t, k, c = r.pos[0], r.idx[0], r.coef[0]
    rt, rk, rc = z[name + '_trace_t'], z[name + '_trace_k'], z[name + '_trace_c']
    cm = TraceComparison(rt, rk, rc, t, k, c)
    n = cm.common_prefix
    err = 0.0
    if n:
        scale = np.maximum(np.abs(rc[:n]), 1e-2 * np.max(np.abs(rc[:n])))
        err = float(np.max(np.abs(rc[:n] - c[:n]) / scale))
    print('%-28s kw=%s prefix=%d ref=%d got=%d coef_err=%.2e snr ref %.3f got %.3f stop=%s' % (name, kw, n, cm.n_ref, cm.n_got, err, snr_db(x, z[name + '_res']), snr_db(x, res), r.stats(0)['stop']))
    if n < min(cm.n_ref, cm.n_got):
        lo = max(0, n - 2)
        print('   ref:', list(zip(rt[lo:n + 4].tolist(), rk[lo:n + 4].tolist(), np.round(rc[lo:n + 4], 6).tolist())))
        print('   got:', list(zip(t[lo:n + 4].tolist(), k[lo:n + 4].tolist(), np.round(c[lo:n + 4], 6).tolist())))
