"""Latency of the single-sequence configs (BASELINE configs 1 and 3, golden fixtures) through the public API, next to the
oracle port on the host: these are latency-bound (one CTA per sequence), the throughput configs are in bench.py."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import hierarchical_sparse_coding_b200 as hsc          # noqa: E402
from oracle import hsc_oracle as O                      # noqa: E402
from helpers import load_npz, case_kwargs                # noqa: E402


def med(fn, n=15):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))


z = load_npz('c1_toy.npz')
for name, cls, ofn in (('c1_cmp', hsc.ConvolutionalMatchingPursuit, O.mp_encode), ('c1_locomp', hsc.LoCOMP, O.locomp_encode)):
    x, D, kw = z[name + '_x'], z[name + '_D'], case_kwargs(z, name)
    coder = hsc.ConvolutionalSparseCoder(D, cls())
    coder.encode(x, **kw)
    g = med(lambda: coder.encode(x, **kw))
    c = med(lambda: ofn(x, D, **kw), n=3)
    n_atoms = len(z[name + '_trace_t'])
    print('%-10s T=%d K=%d L=%d  %d atoms: engine %.2f ms (%.1f us/atom, %.0f atoms/s)   oracle port %.0f ms (%.0fx)' % (
        name, x.shape[0], D.shape[0], D.shape[1], n_atoms, g, 1e3 * g / n_atoms, n_atoms / g * 1e3, c, c / g))

z = np.load(os.path.join(ROOT, 'tests', 'golden', 'c3_complex.npz'))
nl = int(z['nb_levels'])
raw = [z['raw_l%d' % l] for l in range(nl)]
rep = [z['rep_l%d' % l] for l in range(nl)]
cns = z['counts_no_singletons']
scales = [int(v) for v in z['scales']]
mld = hsc.MultilevelDictionary(raw, scales, rep, cns, hasSingletonBases=True)
x = z['x']
for nb in (10, 1):
    coder = hsc.HierarchicalConvolutionalSparseCoder(mld, hsc.HierarchicalConvolutionalMatchingPursuit(method='cmp'))
    codes, res = coder.encode(x, toleranceSnr=10.0, nbBlocks=nb, singletonWeight=0.95)
    g = med(lambda: coder.encode(x, toleranceSnr=10.0, nbBlocks=nb, singletonWeight=0.95), n=7)
    c = med(lambda: O.hierarchical_encode(x, raw, cns, rep, toleranceSnr=10.0, nbBlocks=nb, singletonWeight=0.95), n=1)
    nn = sum(cd.nnz for cd in codes)
    print('c3 hierarchical nbBlocks=%-2d T=%d levels K=%s: %d atoms: engine %.1f ms   oracle port %.0f ms (%.0fx)' % (
        nb, x.shape[0], [d.shape[0] for d in raw], nn, g, c, c / g))
