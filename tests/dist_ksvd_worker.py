"""Worker of tests/test_parity_gpu.py::test_distributed_ksvd_equals_single_process (2 ranks, one GPU each):
data-parallel K-SVD - every rank encodes its own segments, the dictionary update all-reduces one Gram matrix per
filter - must return the dictionary a single process learns on all the segments."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hierarchical_sparse_coding_b200 as hsc          # noqa: E402
from oracle import hsc_oracle as O                      # noqa: E402  (test infrastructure: normalisation only)


def main():
    rank = int(os.environ['RANK'])
    world = int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    dist.init_process_group('nccl')
    rs = np.random.RandomState(17)
    K, L, F, S, T, n = 6, 12, 2, 8, 1500, 40
    Dt = O.normalize(rs.randn(K, L, F))
    x = np.zeros((S, T, F))
    for s in range(S):
        for p, k, a in zip(rs.randint(0, T - L, n), rs.randint(0, K, n), rs.uniform(0.5, 2.0, n)):
            x[s, p:p + L] += a * Dt[k]
    D0 = O.normalize(Dt + 0.4 * rs.randn(K, L, F))
    lo, hi = rank * S // world, (rank + 1) * S // world
    kw = dict(method='cmp', maxIterations=3, toleranceSnr=None, nbNonzeroCoefs=n)
    D_dist = hsc.ConvolutionalDictionaryLearner(K, L, algorithm='ksvd').train(x[lo:hi], initD=D0 if rank == 0 else np.zeros_like(D0),
                                                                              group=True, **kw)
    # every rank holds the same dictionary
    Dg = [torch.zeros(D_dist.shape, dtype=torch.float64, device='cuda') for _ in range(world)]
    dist.all_gather(Dg, torch.from_numpy(D_dist).cuda())
    for d in Dg:
        assert torch.equal(d, Dg[0]), 'ranks disagree on the dictionary'
    if rank == 0:
        D_single = hsc.ConvolutionalDictionaryLearner(K, L, algorithm='ksvd').train(x, initD=D0, **kw)
        err = float(np.max(np.abs(D_dist - D_single)))
        assert err < 1e-9, err
        print('distributed K-SVD == single process, max |dD| = %.2e' % err)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
