"""Generates the committed golden fixtures by running the UNMODIFIED reference.

Run in the dev container only (needs /root/reference):

    python tests/golden/make_golden.py            # all fixtures (~3-4 min, datasets dominate)
    python tests/golden/make_golden.py --skip-datasets

Outputs (all small, committed):
    tests/golden/utils_vectors.npz      peek / overlapAdd / overlapReplace (hsc/utils.py:76-161)
    tests/golden/select_atoms.npz       _selectBestAtoms (hsc/modeling.py:899-982)
    tests/golden/correlate.npz          convolve1d (hsc/modeling.py:149-188)
    tests/golden/mp_cases.npz           MP / LoCOMP traces on seeded inputs (hsc/modeling.py:1053-1425)
    tests/golden/c1_toy.npz             BASELINE config 1 (scripts/demo_csc.py restated)
    tests/golden/c3_complex.npz         BASELINE config 3 (scripts/learn_mlcsc_dataset.py restated)
    tests/golden/ksvd_update.npz        one K-SVD dictionary-update stage (hsc/modeling.py:593-636)
    tests/golden/ksvd_update_pca.npz    the same stage with usePCA=True (:618-625)

Every array the reference consumed is stored next to what it produced, so the tests never need
the reference or its RNG stream again.
"""
import argparse
import os
import sys
import time
import logging

import numpy as np
import scipy.linalg
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

hsc = ref_loader.load_reference()
from hsc.modeling import (ConvolutionalMatchingPursuit, LoCOMP, ConvolutionalSparseCoder,  # noqa: E402
                          HierarchicalConvolutionalMatchingPursuit, HierarchicalConvolutionalSparseCoder,
                          ConvolutionalDictionaryLearner, reconstructSignal, convolve1d)
from hsc.utils import normalize, peek, overlapAdd, overlapReplace  # noqa: E402
from hsc.dataset import MultilevelDictionaryGenerator, SignalGenerator, MultilevelDictionary  # noqa: E402

logging.getLogger('hsc').setLevel(logging.ERROR)


def record_trace(approx):
    """Wraps the approximator's _updateCoefficients so every applied atom is logged in order."""
    log = []
    orig = approx._updateCoefficients

    def wrapped(coefficients, atoms, replace=True):
        for a in atoms:
            log.append((int(a.position), int(a.index), float(a.coefficient)))
        return orig(coefficients, atoms, replace=replace)
    approx._updateCoefficients = wrapped
    return log


def coo_triplets(m):
    c = scipy.sparse.coo_matrix(m)
    order = np.lexsort((c.col, c.row))
    return c.row[order].astype(np.int64), c.col[order].astype(np.int64), c.data[order].astype(np.float64)


def gen_utils(out):
    d = {}
    n = 0
    for T in (10, 11):
        base = np.arange(T, dtype=np.float64) + 1.0
        for width in (1, 2, 3, 4, 5, 6):
            elem = (np.arange(width, dtype=np.float64) + 1.0) * 10.0
            for t in list(range(-8, T + 8)) + [-20]:
                d['case%d_in' % n] = np.array([T, width, t])
                d['case%d_peek' % n] = np.asarray(peek(base, width, t))
                d['case%d_add' % n] = overlapAdd(base, elem, t, copy=True)
                d['case%d_rep' % n] = overlapReplace(base, elem, t, copy=True)
                n += 1
    # 2-D signal
    base2 = np.arange(24, dtype=np.float64).reshape(8, 3)
    for width in (3, 4):
        elem2 = -np.arange(width * 3, dtype=np.float64).reshape(width, 3)
        for t in range(-3, 11):
            d['case%d_in' % n] = np.array([8, width, t])
            d['case%d_peek' % n] = np.asarray(peek(base2, width, t))
            d['case%d_add' % n] = overlapAdd(base2, elem2, t, copy=True)
            d['case%d_rep' % n] = overlapReplace(base2, elem2, t, copy=True)
            n += 1
    d['count'] = np.array(n)
    np.savez_compressed(os.path.join(out, 'utils_vectors.npz'), **d)
    print('utils vectors: %d cases' % n)


def gen_select(out):
    rs = np.random.RandomState(1234)
    cmp = ConvolutionalMatchingPursuit()
    d = {}
    n = 0

    def add(inner, L, nb, offset, thres, weights):
        nonlocal n
        atoms = cmp._selectBestAtoms(inner, L, nb, offset=offset, nullCoeffThres=thres, weights=weights)
        d['s%d_inner' % n] = inner
        d['s%d_par' % n] = np.array([L, -1 if nb == 'auto' else nb, int(offset)], dtype=np.int64)
        d['s%d_thres' % n] = np.array(thres)
        d['s%d_w' % n] = np.zeros(0) if weights is None else weights
        d['s%d_t' % n] = np.array([a.position for a in atoms], dtype=np.int64)
        d['s%d_k' % n] = np.array([a.index for a in atoms], dtype=np.int64)
        d['s%d_c' % n] = np.array([a.coefficient for a in atoms], dtype=np.float64)
        n += 1

    # the reference's own golden inputs (tests/hsc/test_modeling.py:272-325)
    ramp = np.arange(256).reshape((64, 4)).astype(np.float64)
    ramp[-1] = ramp[-1][::-1]
    for (L, nb, off) in ((5, 4, False), (5, 4, True), (3, 'auto', False), (5, 5, False), (5, 5, True),
                         (3, 'auto', True), (5, 1, False)):
        add(ramp, L, nb, off, 0.0, None)
    # random maps, with weights / thresholds / ties
    for T, K, L in ((64, 4, 5), (100, 7, 6), (257, 3, 9), (33, 5, 4), (300, 8, 32)):
        for nb in (1, 2, 3, 8, 'auto'):
            for off in (False, True):
                inner = rs.randn(T, K)
                w = None if rs.rand() < 0.5 else np.where(np.arange(K) < K // 2, 0.5, 1.0)
                add(inner, L, nb, off, 1e-16, w)
    tie = np.zeros((40, 3))
    tie[7, 2] = -2.0
    tie[7, 1] = 2.0
    tie[30, 0] = 2.0
    add(tie, 5, 1, False, 1e-16, None)
    add(tie, 5, 4, False, 1e-16, None)
    add(np.zeros((40, 3)), 5, 1, False, 1e-16, None)
    add(np.zeros((40, 3)), 5, 4, True, 1e-16, None)
    d['count'] = np.array(n)
    np.savez_compressed(os.path.join(out, 'select_atoms.npz'), **d)
    print('select_atoms: %d cases' % n)


def gen_correlate(out):
    rs = np.random.RandomState(99)
    d = {}
    n = 0
    for dtype in (np.float64, np.float32):
        for (T, K, L, F) in ((32, 3, 5, 1), (32, 3, 6, 1), (40, 4, 5, 2), (40, 4, 6, 5), (64, 8, 16, 4),
                             (19, 2, 7, 3), (128, 16, 32, 1)):
            x = rs.randn(T, F).astype(dtype)
            D = normalize(rs.randn(K, L, F)).astype(dtype)
            if F == 1 and n % 2 == 0:
                x1, D1 = x[:, 0], D[:, :, 0]
            else:
                x1, D1 = x, D
            d['c%d_x' % n] = x1
            d['c%d_D' % n] = D1
            d['c%d_same' % n] = convolve1d(x1, D1, padding='same')
            d['c%d_valid' % n] = convolve1d(x1, D1, padding='valid')
            n += 1
    d['count'] = np.array(n)
    np.savez_compressed(os.path.join(out, 'correlate.npz'), **d)
    print('correlate: %d cases' % n)


def _run_case(d, name, method, x, D, kwargs):
    approx = ConvolutionalMatchingPursuit() if method == 'cmp' else LoCOMP()
    log = record_trace(approx)
    coef, res = approx.computeCoefficients(x, D, **kwargs)
    r, c, v = coo_triplets(coef)
    d[name + '_x'] = x
    d[name + '_D'] = D
    d[name + '_method'] = np.array(method)
    kw = dict(kwargs)
    d[name + '_kw'] = np.array(repr({k: (v2.tolist() if isinstance(v2, np.ndarray) else v2) for k, v2 in kw.items()}))
    if len(log) > 0:
        lt, lk, lc = zip(*log)
    else:
        lt, lk, lc = [], [], []
    d[name + '_trace_t'] = np.array(lt, dtype=np.int64)
    d[name + '_trace_k'] = np.array(lk, dtype=np.int64)
    d[name + '_trace_c'] = np.array(lc, dtype=np.float64)
    d[name + '_coo_t'] = r
    d[name + '_coo_k'] = c
    d[name + '_coo_v'] = v
    d[name + '_res'] = res
    return len(log), coef.nnz


def gen_mp_cases(out):
    rs = np.random.RandomState(2024)
    d = {}
    names = []

    def case(name, method, x, D, **kw):
        t0 = time.time()
        n, nnz = _run_case(d, name, method, x, D, kw)
        names.append(name)
        print('  %-28s %-6s events=%5d nnz=%5d  %.2fs' % (name, method, n, nnz, time.time() - t0))

    # known-answer test of the reference (tests/hsc/test_modeling.py:379-396)
    for dtype, tag in ((np.float64, 'f64'), (np.float32, 'f32')):
        D = normalize(rs.random_sample(size=(4, 32)), axis=1).astype(dtype)
        ref = scipy.sparse.coo_matrix(([1.0, 1.0, 0.5, 1.0, 0.75, 2.0], ([32, 48, 64, 96, 128, 192], [0, 3, 1, 0, 2, 2])),
                                      shape=(256, 4))
        x = reconstructSignal(ref, D).astype(dtype)
        case('planted_' + tag, 'cmp', x, D, nbNonzeroCoefs=8, minCoefficients=1e-6)
        case('planted_locomp_' + tag, 'locomp', x, D, minCoefficients=1e-10)

    # property-test shapes of the reference (tests/hsc/test_modeling.py:247-270, :329-377)
    for dtype, tag in ((np.float64, 'f64'), (np.float32, 'f32')):
        for (K, L) in ((1, 3), (2, 5), (3, 6)):
            x = rs.random_sample(size=(16,)).astype(dtype)
            D = normalize(rs.random_sample(size=(K, L)), axis=1).astype(dtype)
            case('tiny_k%d_l%d_%s' % (K, L, tag), 'cmp', x, D, nbNonzeroCoefs=4)
        x = rs.random_sample(size=(64, 7)).astype(dtype)
        D = normalize(rs.random_sample(size=(16, 15, 7)), axis=(1, 2)).astype(dtype)
        case('f7_nnz16_' + tag, 'cmp', x, D, nbNonzeroCoefs=16)
        case('f7_nnz16_locomp_' + tag, 'locomp', x, D, nbNonzeroCoefs=16)
        x = rs.random_sample(size=(128,)).astype(dtype)
        D = normalize(rs.random_sample(size=(32, 9)), axis=1).astype(dtype)
        case('scale0.1_' + tag, 'cmp', x, D, toleranceResidualScale=0.1)
        case('snr10_' + tag, 'cmp', x, D, toleranceSnr=10)
        case('snr20_' + tag, 'cmp', x, D, toleranceSnr=20)
        for F in (4, 11):
            x = rs.random_sample(size=(128, F)).astype(dtype)
            D = normalize(rs.random_sample(size=(32, 9, F)), axis=(1, 2)).astype(dtype)
            case('feat%d_scale_%s' % (F, tag), 'cmp', x, D, toleranceResidualScale=0.05 if F == 4 else 0.2)

    # zero-mean signals (sign handling), weights, even/odd lengths, many edge atoms
    for dtype, tag in ((np.float64, 'f64'), (np.float32, 'f32')):
        for (T, K, L, F) in ((96, 6, 8, 1), (96, 6, 9, 2), (50, 5, 16, 3), (40, 3, 12, 1)):
            x = rs.randn(T, F).astype(dtype)
            D = normalize(rs.randn(K, L, F)).astype(dtype)
            if F == 1:
                x, D = x[:, 0], D[:, :, 0]
            case('randn_T%d_L%d_F%d_%s' % (T, L, F, tag), 'cmp', x, D, nbNonzeroCoefs=40)
            w = np.where(np.arange(K) < K // 2, 0.5, 1.0).astype(dtype)
            case('weights_T%d_L%d_F%d_%s' % (T, L, F, tag), 'cmp', x, D, toleranceSnr=8.0, weights=w)

    # block selection (tests/hsc/test_modeling.py:209-215 shapes)
    for dtype, tag in ((np.float64, 'f64'), (np.float32, 'f32')):
        x = rs.random_sample(size=(256,)).astype(dtype)
        D = normalize(rs.random_sample(size=(16, 15)), axis=1).astype(dtype)
        for nb in (2, 8, 'auto'):
            case('blocks%s_%s' % (nb, tag), 'cmp', x, D, toleranceSnr=5.0, nbBlocks=nb)
            case('blocks%s_locomp_%s' % (nb, tag), 'locomp', x, D, toleranceSnr=5.0, nbBlocks=nb)
        x = rs.randn(512, 3).astype(dtype)
        D = normalize(rs.randn(8, 10, 3)).astype(dtype)
        case('blocks10_randn_' + tag, 'cmp', x, D, toleranceSnr=6.0, nbBlocks=10)
        case('blocks10_nnz_' + tag, 'cmp', x, D, nbNonzeroCoefs=60, nbBlocks=10)

    # planted overlapping atoms, sparse support (the benchmark signal law, small)
    for dtype, tag in ((np.float64, 'f64'), (np.float32, 'f32')):
        T, K, L, F = 2048, 16, 32, 4
        D = normalize(rs.randn(K, L, F)).astype(dtype)
        n = 40
        ref = scipy.sparse.coo_matrix((rs.uniform(0.25, 4.0, n) * rs.choice([-1.0, 1.0], n),
                                       (rs.randint(0, T, n), rs.randint(0, K, n))), shape=(T, K))
        x = reconstructSignal(ref.tocsc(), D).astype(dtype)
        case('sparse2048_' + tag, 'cmp', x, D, toleranceSnr=40.0)
        case('sparse2048_locomp_' + tag, 'locomp', x, D, toleranceSnr=40.0)

    d['names'] = np.array(names)
    np.savez_compressed(os.path.join(out, 'mp_cases.npz'), **d)
    print('mp cases: %d' % len(names))


def _consume_visualize_rng(mld):
    """The generator scripts call multilevelDict.visualize(maxCounts=16) between the dictionary and the
    signals (scripts/generate_dataset.py:74); with shuffle=True it draws one np.random.permutation per
    level (hsc/dataset.py:356-358).  Drawn here too so the signal RNG stream is the scripts' own."""
    for level in range(mld.getNbLevels()):
        np.random.permutation(np.arange(mld.representations[level].shape[0], dtype=int))


def _generate_toy():
    """scripts/generate_dataset_toy.py:95-139 restated (same RNG stream: dict, train 1e7, test 1e5)."""
    import importlib.util
    # the script defines ToySignalGenerator; load it through the py2 loader without running __main__
    path = os.path.join(ref_loader.REFERENCE_ROOT, 'scripts', 'generate_dataset_toy.py')
    loader = ref_loader._Py2Loader('hsc_script_toy', path)
    spec = importlib.util.spec_from_loader('hsc_script_toy', loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    np.random.seed(42)
    mld = MultilevelDictionaryGenerator().generate([16, 64, 128, 256], [4, 8, 16, 32], decompositionSize=[None, 3, 2, 2],
                                                   positionSampling='no-overlap', weightSampling='random',
                                                   multilevelDecomposition=False, maxNbPatternsConsecutiveRejected=1000,
                                                   nonNegativity=False)
    _consume_visualize_rng(mld)
    sig = {}
    for name, n in (('train', int(1e7)), ('test', int(1e5))):
        gen = mod.ToySignalGenerator(mld, rates=[None] * mld.getNbLevels())
        ev = gen.generateEvents(n)
        sig[name] = gen.generateSignalFromEvents(ev, nbSamples=n)
        print('  toy %s: %d events' % (name, len(ev)))
    return mld, sig


def _generate_complex():
    """scripts/generate_dataset.py:46-94 restated."""
    np.random.seed(42)
    scales = [32, 64, 128, 256]
    mld = MultilevelDictionaryGenerator().generate(scales, [4, 8, 16, 32], decompositionSize=[None, 4, 3, 3],
                                                   positionSampling='random', weightSampling='random',
                                                   multilevelDecomposition=False, maxNbPatternsConsecutiveRejected=1000,
                                                   nonNegativity=False)
    _consume_visualize_rng(mld)
    sig = {}
    rates_init = 0.0005 * np.ones_like(scales)
    for name, n in (('train', int(1e7)), ('test', int(1e5))):
        gen = SignalGenerator(mld, rates_init)
        ev, rates = gen.generateEvents(n, 0.25)
        sig[name] = gen.generateSignalFromEvents(ev, nbSamples=n)
        print('  complex %s: %d events' % (name, len(ev)))
    return mld, sig


def gen_c1(out):
    t0 = time.time()
    mld, sig = _generate_toy()
    print('  toy dataset generated in %.1fs' % (time.time() - t0))
    D = mld.dictionaries[0]
    x = sig['test'][:10000]
    d = {}
    for method in ('cmp', 'locomp'):
        t0 = time.time()
        n, nnz = _run_case(d, 'c1_' + method, method, x, D, dict(toleranceSnr=20.0, nbBlocks=1))
        print('  c1 %s: events=%d nnz=%d %.2fs' % (method, n, nnz, time.time() - t0))
    # a slice of the training signal for the C2-shaped parity case (first 50k samples, 16 sampled filters)
    np.random.seed(42)
    D16 = ConvolutionalDictionaryLearner(k=16, windowSize=32, algorithm='samples').train(sig['train'][:10000])
    xs = sig['train'][:50000]
    _run_case(d, 'c2s_cmp', 'cmp', xs, D16, dict(nbNonzeroCoefs=150))
    print('  c2 slice: events=%d' % len(d['c2s_cmp_trace_t']))
    np.savez_compressed(os.path.join(out, 'c1_toy.npz'), **d)


def gen_c3(out):
    t0 = time.time()
    mld, sig = _generate_complex()
    print('  complex dataset generated in %.1fs' % (time.time() - t0))
    mld2 = mld.upToLevel(2)
    x = sig['test'][:20000]
    d = {'x': x}
    coder_mld = None
    for tag, method, nb in (('cmp_b10', 'cmp', 10), ('cmp_b1', 'cmp', 1), ('locomp_b10', 'locomp', 10)):
        hcmp = HierarchicalConvolutionalMatchingPursuit(method=method)
        hcsc = HierarchicalConvolutionalSparseCoder(mld2, approximator=hcmp)
        coder_mld = hcsc.multilevelDict
        t0 = time.time()
        for dist in (True, False):
            codes, res = hcsc.encode(x, toleranceSnr=10.0, nbBlocks=nb, singletonWeight=0.95, returnDistributed=dist)
            sfx = '' if dist else '_nodist'
            for lvl, c in enumerate(codes):
                r, cc, v = coo_triplets(c)
                d['%s%s_l%d_t' % (tag, sfx, lvl)] = r
                d['%s%s_l%d_k' % (tag, sfx, lvl)] = cc
                d['%s%s_l%d_v' % (tag, sfx, lvl)] = v
            d['%s%s_res' % (tag, sfx)] = res
        print('  c3 %s: nnz/level=%s  %.2fs' % (tag, [c.nnz for c in codes], time.time() - t0))
    nl = coder_mld.getNbLevels()
    d['nb_levels'] = np.array(nl)
    d['counts_no_singletons'] = np.asarray(coder_mld.countsNoSingletons)
    d['scales'] = np.asarray(coder_mld.scales)
    for lvl in range(nl):
        d['raw_l%d' % lvl] = coder_mld.getRawDictionary(lvl)
        d['rep_l%d' % lvl] = coder_mld.getMultiscaleDictionaries()[lvl]
    np.savez_compressed(os.path.join(out, 'c3_complex.npz'), **d)


def gen_ksvd(out):
    rs = np.random.RandomState(7)
    d = {}
    n = 0
    for (T, K, L, F) in ((400, 6, 8, 1), (300, 5, 9, 3)):
        Dt = normalize(rs.randn(K, L, F))
        D0 = normalize(rs.randn(K, L, F))
        ref = scipy.sparse.coo_matrix((rs.uniform(0.5, 2.0, 30), (rs.randint(L, T - L, 30), rs.randint(0, K, 30))), shape=(T, K))
        if F == 1:
            Dt, D0 = Dt[:, :, 0], D0[:, :, 0]
        x = reconstructSignal(ref.tocsc(), Dt)
        coef, res = ConvolutionalSparseCoder(D0, ConvolutionalMatchingPursuit()).encode(x, nbNonzeroCoefs=25)
        # one dictionary-update stage, replayed with the reference's own statements (hsc/modeling.py:594-636)
        from hsc.modeling import extractWindows
        D = np.copy(D0)
        coefficients = coef.copy()
        for k in range(D.shape[0]):
            indices = coefficients[:, k].nonzero()[0]
            if len(indices) == 0:
                continue
            coefficients[indices, k * np.ones_like(indices)] = 0.0
            error = reconstructSignal(coefficients, D)
            windows = extractWindows(np.pad(error, [(D.shape[1] // 2, D.shape[1] // 2), ] + [(0, 0) for _ in range(error.ndim - 1)], mode='constant'),
                                     D.shape[1] // 2 + indices, width=D.shape[1], centered=True)
            windows = windows.reshape((windows.shape[0], -1))
            U, s, Vh = scipy.linalg.svd(windows.T, full_matrices=False)
            D[k, :] = U[:, 0].reshape(D.shape[1:])
            coefficients[indices, k * np.ones_like(indices)] = Vh.T[:, 0] * s[0]
        d['k%d_x' % n] = x
        d['k%d_D0' % n] = D0
        r, c, v = coo_triplets(coef)
        d['k%d_code_t' % n], d['k%d_code_k' % n], d['k%d_code_v' % n] = r, c, v
        d['k%d_D1' % n] = D
        r, c, v = coo_triplets(coefficients)
        d['k%d_code1_t' % n], d['k%d_code1_k' % n], d['k%d_code1_v' % n] = r, c, v
        n += 1
    d['count'] = np.array(n)
    np.savez_compressed(os.path.join(out, 'ksvd_update.npz'), **d)
    print('ksvd cases: %d' % n)


def gen_ksvd_pca(out):
    """The usePCA=True branch of the dictionary-update stage (hsc/modeling.py:618-625), replayed with the reference's
    own statements and its own pca() (:48-80)."""
    from hsc.modeling import extractWindows, pca
    rs = np.random.RandomState(17)
    d = {}
    n = 0
    for (T, K, L, F) in ((400, 6, 8, 1), (300, 5, 9, 3)):
        Dt = normalize(rs.randn(K, L, F))
        D0 = normalize(rs.randn(K, L, F))
        ref = scipy.sparse.coo_matrix((rs.uniform(0.5, 2.0, 30), (rs.randint(L, T - L, 30), rs.randint(0, K, 30))), shape=(T, K))
        if F == 1:
            Dt, D0 = Dt[:, :, 0], D0[:, :, 0]
        x = reconstructSignal(ref.tocsc(), Dt)
        coef, res = ConvolutionalSparseCoder(D0, ConvolutionalMatchingPursuit()).encode(x, nbNonzeroCoefs=25)
        D = np.copy(D0)
        coefficients = coef.copy()
        for k in range(D.shape[0]):
            indices = coefficients[:, k].nonzero()[0]
            if len(indices) == 0:
                continue
            coefficients[indices, k * np.ones_like(indices)] = 0.0
            error = reconstructSignal(coefficients, D)
            windows = extractWindows(np.pad(error, [(D.shape[1] // 2, D.shape[1] // 2), ] + [(0, 0) for _ in range(error.ndim - 1)], mode='constant'),
                                     D.shape[1] // 2 + indices, width=D.shape[1], centered=True)
            windows = windows.reshape((windows.shape[0], -1))
            evals, evecs = pca(windows, k=1)
            evec = evecs[:, 0]
            D[k, :] = evec.reshape(D.shape[1:])
            coefficients[indices, k * np.ones_like(indices)] = np.dot(windows, evec[:, np.newaxis])[:, 0]
        d['k%d_x' % n] = x
        d['k%d_D0' % n] = D0
        r, c, v = coo_triplets(coef)
        d['k%d_code_t' % n], d['k%d_code_k' % n], d['k%d_code_v' % n] = r, c, v
        d['k%d_D1' % n] = D
        r, c, v = coo_triplets(coefficients)
        d['k%d_code1_t' % n], d['k%d_code1_k' % n], d['k%d_code1_v' % n] = r, c, v
        n += 1
    d['count'] = np.array(n)
    np.savez_compressed(os.path.join(out, 'ksvd_update_pca.npz'), **d)
    print('ksvd pca cases: %d' % n)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--skip-datasets', action='store_true')
    ap.add_argument('--only', default=None)
    args = ap.parse_args()
    import warnings
    warnings.simplefilter('ignore')
    steps = [('utils', gen_utils), ('select', gen_select), ('correlate', gen_correlate), ('mp', gen_mp_cases),
             ('ksvd', gen_ksvd), ('ksvd_pca', gen_ksvd_pca)]
    if not args.skip_datasets:
        steps += [('c1', gen_c1), ('c3', gen_c3)]
    for name, fn in steps:
        if args.only and name not in args.only.split(','):
            continue
        print('== %s' % name)
        fn(HERE)
