"""Golden vectors of the convolutional k-means learner and of the event <-> sparse converters, recorded from the
UNMODIFIED reference (dev container only):

    python tests/golden/make_golden_kmeans.py

    tests/golden/kmeans.npz        ConvolutionalDictionaryLearner(algorithm='kmean').train (hsc/modeling.py:420-526):
                                   data, np.random seed, arguments -> dictionary after 1 and after N iterations, for the
                                   three reset methods / two init methods, 1-D and multichannel, INCLUDING a case where
                                   a centroid's only window is window 0 (the reference's empty test, :479-481, resets it)
    tests/golden/converters.npz    convertSparseMatricesToEvents / convertEventsToSparseMatrices (hsc/dataset.py:798-824)
"""
import os
import sys
import logging

import numpy as np
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

hsc = ref_loader.load_reference()
from hsc.modeling import ConvolutionalDictionaryLearner, convolve1d_batch, extractRandomWindows  # noqa: E402
from hsc.dataset import convertSparseMatricesToEvents, convertEventsToSparseMatrices  # noqa: E402

logging.getLogger('hsc').setLevel(logging.ERROR)


def smooth_data(rs, T, F, dtype):
    d = rs.randn(T, F)
    base = np.convolve(d[:, 0], np.hanning(7), 'same')[:, None] * np.ones((1, F)) + 0.1 * d
    base = base.astype(dtype)
    return base[:, 0] if F == 1 else base


def only_window0_case(data, k, W, nb, seed):
    """True if, in the FIRST iteration under this seed, some centroid's only window is window 0."""
    np.random.seed(seed)
    windows = extractRandomWindows(data, nb, 2 * W)
    D = ConvolutionalDictionaryLearner(k, W, algorithm='kmean')._init_D(data, 'random_samples')
    ip = convolve1d_batch(windows, D, padding='valid')
    idx = np.unravel_index(np.argmax(np.abs(ip.reshape(ip.shape[0], -1)), axis=1), ip.shape[1:])[1]
    return np.sum(idx == idx[0]) == 1


def gen_kmeans(out):
    rs = np.random.RandomState(31)
    d = {}
    n = 0

    def case(data, k, W, nb, seed, iters, **kw):
        nonlocal n
        for it in (1, iters):
            np.random.seed(seed)
            D = ConvolutionalDictionaryLearner(k, W, algorithm='kmean').train(data, nbRandomWindows=nb, maxIterations=it, **kw)
            d['m%d_D_it%d' % (n, it)] = D
        d['m%d_data' % n] = data
        d['m%d_par' % n] = np.array([k, W, nb, seed, iters])
        d['m%d_kw' % n] = np.array(repr(kw))
        n += 1

    for (T, F, k, W, nb, dtype) in ((5000, 1, 6, 16, 400, np.float32), (4000, 3, 5, 9, 300, np.float32), (3000, 2, 8, 8, 200, np.float64)):
        data = smooth_data(rs, T, F, dtype)
        case(data, k, W, nb, 11, 4)
        case(data, k, W, nb, 12, 3, resetMethod='random_samples')
        case(data, k, W, nb, 13, 3, initMethod='noise', resetMethod='random_samples_average', nbAveragedPatches=4)
    # many centroids, few windows: empty centroids in every iteration, and a seed where window 0 is alone in its cluster
    data = smooth_data(rs, 2000, 1, np.float32)
    seed = next(s for s in range(1000) if only_window0_case(data, 16, 8, 24, s))
    for reset in ('noise', 'random_samples', 'random_samples_average'):
        case(data, 16, 8, 24, seed, 3, resetMethod=reset)
    d['window0_case_first'] = np.array(n - 3)
    d['count'] = np.array(n)
    np.savez_compressed(os.path.join(out, 'kmeans.npz'), **d)
    print('kmeans cases: %d (window-0 seed %d)' % (n, seed))


def gen_converters(out):
    rs = np.random.RandomState(5)
    d = {}
    T, counts = 300, [4, 7, 9]
    codes = []
    for lv, K in enumerate(counts):
        dense = rs.randn(T, K) * (rs.rand(T, K) < 0.04)
        codes.append(scipy.sparse.csr_matrix(dense.astype(np.float32)))
        c = codes[-1].tocoo()
        d['code%d_t' % lv], d['code%d_k' % lv], d['code%d_v' % lv] = c.row, c.col, c.data
    ev = convertSparseMatricesToEvents(codes)
    d['events'] = np.stack([ev['f0'], ev['f1'], ev['f2']], axis=1)
    d['events_v'] = ev['f3']
    back = convertEventsToSparseMatrices(ev, counts, T)
    for lv, m in enumerate(back):
        assert m.format == 'csr'
        c = m.tocoo()
        d['back%d_t' % lv], d['back%d_k' % lv], d['back%d_v' % lv] = c.row, c.col, c.data
    d['counts'] = np.array(counts)
    d['T'] = np.array(T)
    np.savez_compressed(os.path.join(out, 'converters.npz'), **d)
    print('converters: %d events' % len(ev))


if __name__ == '__main__':
    import warnings
    warnings.simplefilter('ignore')
    gen_kmeans(HERE)
    gen_converters(HERE)
