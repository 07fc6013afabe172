import os, sys, cProfile, pstats, io
import numpy as np
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import hierarchical_sparse_coding_b200 as hsc
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'c3_complex.npz'))
nl = int(z['nb_levels'])
raw = [z['raw_l%d' % l] for l in range(nl)]; rep = [z['rep_l%d' % l] for l in range(nl)]
mld = hsc.MultilevelDictionary(raw, [int(v) for v in z['scales']], rep, z['counts_no_singletons'], hasSingletonBases=True)
x = z['x']
coder = hsc.HierarchicalConvolutionalSparseCoder(mld, hsc.HierarchicalConvolutionalMatchingPursuit(method='cmp'))
for _ in range(3): coder.encode(x, toleranceSnr=10.0, nbBlocks=1, singletonWeight=0.95)
pr = cProfile.Profile(); pr.enable()
for _ in range(10): coder.encode(x, toleranceSnr=10.0, nbBlocks=1, singletonWeight=0.95)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:5000])
