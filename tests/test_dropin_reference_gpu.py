"""Drop-in proof: the REFERENCE's own classes drive the B200 approximator.

The reference injects a sparse approximator by composition - ConvolutionalSparseCoder(D, approximator)
(hsc/modeling.py:1656-1669), HierarchicalConvolutionalSparseCoder(multilevelDict, approximator) (:1671-1705) - and its
K-SVD learner constructs one by name (:580-589).  Here the unmodified reference (baseline/_ref on the GPU box, read
through the Python-2 import hook tests/golden/ref_loader.py) is given this package's approximators and must return what
it returns with its own NumPy approximators.
"""
import os
import sys

import numpy as np
import pytest
import scipy.sparse

from helpers import load_npz, snr_db, code_diff

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
import ref_loader  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ref():
    if not ref_loader.reference_available():
        pytest.skip('no copy of the reference (baseline/_ref: python baseline/install_reference.py)')
    import logging
    import warnings
    warnings.simplefilter('ignore')
    mod = ref_loader.load_reference()
    logging.getLogger('hsc').setLevel(logging.ERROR)
    return mod


@pytest.fixture(scope='module')
def hsc():
    import torch
    assert torch.cuda.is_available()
    import hierarchical_sparse_coding_b200 as pkg
    return pkg


def test_reference_coder_with_b200_approximator(ref, hsc):
    """ConvolutionalSparseCoder of the reference, approximator = B200 engine (:1656-1669): encode + reconstruct."""
    rs = np.random.RandomState(3)
    for (T, K, L, F, n, dtype) in ((3000, 12, 16, 1, 60, np.float32), (2048, 16, 32, 4, 40, np.float32), (1500, 8, 9, 3, 30, np.float64)):
        D = ref.utils.normalize(rs.randn(K, L, F)).astype(dtype)
        planted = scipy.sparse.coo_matrix((rs.uniform(0.25, 4.0, n) * rs.choice([-1.0, 1.0], n), (rs.randint(0, T, n), rs.randint(0, K, n))),
                                          shape=(T, K)).tocsc()
        if F == 1:
            D = D[:, :, 0]
        x = ref.modeling.reconstructSignal(planted, D).astype(dtype)
        for kw in (dict(nbNonzeroCoefs=n), dict(toleranceSnr=25.0)):
            theirs = ref.modeling.ConvolutionalSparseCoder(D, approximator=ref.modeling.ConvolutionalMatchingPursuit())
            ours = ref.modeling.ConvolutionalSparseCoder(D, approximator=hsc.ConvolutionalMatchingPursuit())
            c_ref, r_ref = theirs.encode(x, **kw)
            c_got, r_got = ours.encode(x, **kw)
            assert scipy.sparse.issparse(c_got) and c_got.format == c_ref.format and c_got.shape == c_ref.shape and c_got.dtype == c_ref.dtype
            assert r_got.shape == r_ref.shape and r_got.dtype == r_ref.dtype
            ratio, mism = code_diff(c_ref, c_got, rel=1e-5)
            assert mism == 0 and ratio <= 1.0, (T, K, L, F, kw, ratio, mism)
            assert abs(snr_db(x, r_ref) - snr_db(x, r_got)) <= 0.01
            # the reference's decoder on the engine's code (and the engine's decoder on the reference's)
            assert np.allclose(ours.reconstruct(c_got) + r_got, x, atol=1e-5)
            assert np.allclose(hsc.reconstructSignal(c_ref, D), theirs.reconstruct(c_ref), atol=1e-6)
        # LoCOMP through the same composition
        c_ref, r_ref = ref.modeling.ConvolutionalSparseCoder(D, ref.modeling.LoCOMP()).encode(x, nbNonzeroCoefs=n)
        c_got, r_got = ref.modeling.ConvolutionalSparseCoder(D, hsc.LoCOMP()).encode(x, nbNonzeroCoefs=n)
        assert abs(snr_db(x, r_ref) - snr_db(x, r_got)) <= 0.01
        ratio, mism = code_diff(c_ref, c_got, rel=2e-4 if dtype == np.float32 else 1e-6)
        assert mism == 0 and ratio <= 1.0, (T, K, L, F, ratio, mism)


def test_reference_hierarchical_coder_with_b200_approximator(ref, hsc):
    """HierarchicalConvolutionalSparseCoder of the reference over the reference's own MultilevelDictionary
    (hsc/dataset.py:110), approximator = this package's hierarchical MP: config 3's dictionaries and signal."""
    z = load_npz('c3_complex.npz')
    nl = int(z['nb_levels'])
    raw = [z['raw_l%d' % i] for i in range(nl)]
    rep = [z['rep_l%d' % i] for i in range(nl)]
    scales = [int(v) for v in z['scales']]
    mld = ref.dataset.MultilevelDictionary(raw, scales, rep, None, hasSingletonBases=True)
    assert np.array_equal(mld.countsNoSingletons, z['counts_no_singletons'])
    x = z['x'][:6000]
    kw = dict(toleranceSnr=10.0, nbBlocks=1, singletonWeight=0.95, returnDistributed=True)
    theirs = ref.modeling.HierarchicalConvolutionalSparseCoder(mld, ref.modeling.HierarchicalConvolutionalMatchingPursuit(method='cmp'))
    ours = ref.modeling.HierarchicalConvolutionalSparseCoder(mld, hsc.HierarchicalConvolutionalMatchingPursuit(method='cmp'))
    c_ref, r_ref = theirs.encode(x, **kw)
    c_got, r_got = ours.encode(x, **kw)
    assert len(c_got) == len(c_ref) == nl
    assert abs(snr_db(x, r_ref) - snr_db(x, r_got)) <= 0.01
    for l in range(nl):
        assert c_got[l].shape == c_ref[l].shape
        ratio, mism = code_diff(c_ref[l], c_got[l], rel=1e-5)
        assert mism == 0 and ratio <= 1.0, (l, ratio, mism, c_ref[l].nnz, c_got[l].nnz)
    assert np.allclose(ours.reconstruct(c_got) + r_got, x, atol=1e-5)


def test_reference_ksvd_loop_with_b200_inference(ref, hsc, monkeypatch):
    """_train_ksvd of the reference (:528-641) builds its approximator by name (:580-589); with the names bound to this
    package's classes - the one-line change INTEGRATION.md shows - the reference's own loop (its dictionary update, its
    np.random initialisation) learns the same dictionary as with its NumPy approximators."""
    rs = np.random.RandomState(8)
    K, L, T, n = 6, 12, 1500, 60
    Dt = ref.utils.normalize(rs.randn(K, L))
    planted = scipy.sparse.coo_matrix((rs.uniform(0.5, 2.0, n), (rs.randint(0, T, n), rs.randint(0, K, n))), shape=(T, K)).tocsc()
    x = ref.modeling.reconstructSignal(planted, Dt)
    for method in ('cmp', 'locomp'):
        np.random.seed(5)
        D_ref = ref.modeling.ConvolutionalDictionaryLearner(K, L, algorithm='ksvd').train(x, method=method, maxIterations=3, nbNonzeroCoefs=40, toleranceSnr=30.0)
        with monkeypatch.context() as m:
            m.setattr(ref.modeling, 'ConvolutionalMatchingPursuit', hsc.ConvolutionalMatchingPursuit)
            m.setattr(ref.modeling, 'LoCOMP', hsc.LoCOMP)
            np.random.seed(5)
            D_got = ref.modeling.ConvolutionalDictionaryLearner(K, L, algorithm='ksvd').train(x, method=method, maxIterations=3, nbNonzeroCoefs=40, toleranceSnr=30.0)
        assert D_got.shape == D_ref.shape
        # the sign of a singular pair is LAPACK's choice inside the reference's own update (:627-633): compare up to it
        sgn = np.sign(np.sum(D_got * D_ref, axis=1))[:, None]
        assert np.allclose(D_got * sgn, D_ref, atol=1e-6), (method, float(np.abs(D_got * sgn - D_ref).max()))
