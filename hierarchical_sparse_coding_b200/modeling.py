"""Host-side mirror of the reference's encoder/decoder API for the matching-pursuit path
(`hsc.modeling`, all citations relative to the reference tree).  Same class names, keyword
arguments, return types and error behaviour; every computation runs in the CUDA engine behind the
C ABI (include/hsc_b200.h).  There is no CPU fallback: constructing an engine without a CUDA
device raises.

    SparseApproximator                         hsc/modeling.py:657-660
    ConvolutionalMatchingPursuit               hsc/modeling.py:866-1186
    HierarchicalConvolutionalMatchingPursuit   hsc/modeling.py:1427-1654
    ConvolutionalSparseCoder                   hsc/modeling.py:1656-1669
    HierarchicalConvolutionalSparseCoder       hsc/modeling.py:1671-1705
    convolve1d / reconstructSignal             hsc/modeling.py:149-188 / :226-263

New entry points (the reference encodes one signal per call, :1664): `computeCoefficientsBatch` /
`encodeBatch` take [B,T,F] and shard nothing themselves - see `distributed.py` for the multi-GPU
partition by independent signals.
"""
import collections.abc
import copy
import logging

import numpy as np
import scipy.sparse

from . import _native as N
from .engine import Engine, get_engine, engine_dtype

logger = logging.getLogger(__name__)


class MultilevelDictionary(object):
    """Plain holder with the accessors the hierarchical encoder uses from the reference's
    hsc.dataset.MultilevelDictionary (:110-410): raw per-level dictionaries (singleton bases
    already added), input-level representations, countsNoSingletons.  Any object with the same
    attributes/methods (e.g. the reference's own class) is accepted wherever this one is."""

    def __init__(self, dictionaries, scales, representations, countsNoSingletons, hasSingletonBases=True):
        assert len(dictionaries) > 0 and len(scales) == len(dictionaries)
        self.dictionaries = list(dictionaries)
        self.scales = list(scales)
        self.representations = list(representations)
        self.counts = np.array([d.shape[0] for d in dictionaries], dtype=int)
        self.countsNoSingletons = np.asarray(countsNoSingletons, dtype=int)
        self.hasSingletonBases = hasSingletonBases

    def getNbLevels(self):
        return len(self.scales)

    def getRawDictionary(self, level):
        assert level >= 0 and level < self.getNbLevels()
        return self.dictionaries[level]

    def getBaseDictionary(self):
        return self.dictionaries[0]

    def getMultiscaleDictionaries(self):
        return self.representations

    def withSingletonBases(self):
        if self.hasSingletonBases or self.getNbLevels() == 1:
            return self
        raise NotImplementedError('compose the singleton bases with the reference data model '
                                  '(hsc.dataset.addSingletonBases); it is outside the hot path')


def _is_multilevel_dictionary(obj):
    return all(hasattr(obj, a) for a in ('getNbLevels', 'getRawDictionary', 'getMultiscaleDictionaries',
                                         'countsNoSingletons', 'hasSingletonBases'))


def convolve1d(sequence, filters, padding='valid', device=None):
    """hsc/modeling.py:149-188 on the device: [T] or [T,F] x [K,L] or [K,L,F] -> [T',K] ndarray."""
    sequence = np.asarray(sequence)
    filters = np.asarray(filters)
    x = np.atleast_2d(sequence).reshape((sequence.shape[0], -1))
    L = filters.shape[1]
    F = 1 if filters.ndim == 2 else filters.shape[-1]
    assert F == x.shape[-1]
    if padding not in ('valid', 'same'):
        raise Exception('Padding not supported: %s' % (padding))
    eng = get_engine(device)
    dt = engine_dtype(x, filters)
    eng.set_dictionary(filters, dtype=dt)
    c = eng.correlate(x[None].astype(dt)).cpu().numpy()[0]
    if padding == 'valid':
        off = L // 2 - 1 if L % 2 == 0 else L // 2
        c = c[off:off + x.shape[0] - L + 1]
    return c


def reconstructSignal(coefficients, D, device=None):
    """Sparse decoder (hsc/modeling.py:226-263): float64 like the reference (:238)."""
    assert coefficients.ndim == 1 or coefficients.ndim == 2
    D = np.asarray(D)
    assert D.ndim == 2 or D.ndim == 3
    squeeze = D.ndim == 2
    cx = scipy.sparse.coo_matrix(coefficients)
    eng = get_engine(device)
    out_dtype = cx.dtype if cx.dtype in (np.float32, np.float64) else np.float64
    eng.set_dictionary(D, dtype=np.float64)
    keep = cx.data != 0.0
    sig = eng.decode(cx.row[keep], cx.col[keep], cx.data[keep], cx.shape[0]).cpu().numpy().astype(out_dtype)
    return sig[:, 0] if squeeze else sig


class SparseApproximator(object):

    def computeCoefficients(self, X, D):
        raise NotImplementedError()


class ConvolutionalMatchingPursuit(SparseApproximator):
    """Drop-in for hsc.modeling.ConvolutionalMatchingPursuit (:866-1186) on the B200 engine."""

    _method = 0          # hsc_mp_options.method: 0 = MP, 1 = LoCOMP

    def __init__(self, verbose=False, device=None, coef_mode=1):
        self.verbose = verbose
        self.device = device
        self.coef_mode = coef_mode
        self.last_result = None

    def _engine(self):
        return get_engine(self.device)

    def _encode(self, sequences, D, nbNonzeroCoefs, toleranceResidualScale, toleranceSnr, nbBlocks, minCoefficients,
                weights, stopCondition, max_events_total=0):
        eng = self._engine()
        dt = engine_dtype(sequences, D)
        eng.set_dictionary(D, weights=weights, dtype=dt)
        opt = eng.make_options(nbNonzeroCoefs, toleranceResidualScale, toleranceSnr, nbBlocks, minCoefficients,
                               use_weights=weights is not None, coef_mode=self.coef_mode,
                               max_passes_per_run=1 if stopCondition is not None else 0,
                               max_events_total=max_events_total, method=self._method)
        on_pass = None
        if stopCondition is not None:
            x0 = np.asarray(sequences[0])

            def on_pass(res, states):
                if self._method == 1:       # LoCOMP's callback takes the code only (hsc/modeling.py:1395-1396)
                    return bool(stopCondition(res.to_csc(0, None).tolil()))
                # hsc/modeling.py:1155-1158: stopCondition(sequence, residual, coefficients)
                return bool(stopCondition(x0, res.residual[0].cpu().numpy(), res.to_csc(0, None).tolil()))
        res = eng.encode(np.ascontiguousarray(sequences, dtype=dt), opt, on_pass=on_pass)
        self.last_result = res
        if any(st.status == N.HSC_STOP_GROUP for st in res.states):
            raise NotImplementedError('LoCOMP: a selection has more than 63 common-support atoms; the device refit '
                                      'holds groups of at most 64')
        return res

    def computeCoefficients(self, sequence, D, nbNonzeroCoefs=None, toleranceResidualScale=None, toleranceSnr=None,
                            nbBlocks=1, minCoefficients=1e-16, weights=None, stopCondition=None):
        sequence = np.asarray(sequence)
        D = np.asarray(D)
        assert sequence.ndim == 1 or sequence.ndim == 2
        assert D.ndim == 2 or D.ndim == 3
        squeezeOutput = sequence.ndim == 1 or D.ndim == 2
        x = sequence[:, None] if sequence.ndim == 1 else sequence
        if weights is not None:
            assert len(weights) == D.shape[0]
        res = self._encode(x[None], D, nbNonzeroCoefs, toleranceResidualScale, toleranceSnr, nbBlocks, minCoefficients,
                           weights, stopCondition)
        st = res.stats(0)
        logger.debug('SNR stop=%s after %d selection iterations; nnz=%d duplicates=%d', st['stop'], st['passes'],
                     st['nnz'], st['duplicates'])
        if st['stop'] == 'empty':
            logger.warning('Selection returned empty set: considering convergence is achieved')
        coefficients = res.to_csc(0, minCoefficients)
        residual = res.residual[0].cpu().numpy().astype(sequence.dtype if sequence.dtype.kind == 'f' else np.float64)
        if squeezeOutput:
            residual = np.squeeze(residual, axis=1)
        return coefficients, residual

    def computeCoefficientsBatch(self, sequences, D, nbNonzeroCoefs=None, toleranceResidualScale=None, toleranceSnr=None,
                                 nbBlocks=1, minCoefficients=1e-16, weights=None):
        """[B,T] or [B,T,F] -> (list of csc_matrix [T,K], residual ndarray like the input)."""
        sequences = np.asarray(sequences)
        D = np.asarray(D)
        assert sequences.ndim == 2 or sequences.ndim == 3
        assert D.ndim == 2 or D.ndim == 3
        x = sequences[:, :, None] if sequences.ndim == 2 else sequences
        eng = self._engine()
        dt = engine_dtype(x, D)
        eng.set_dictionary(D, weights=weights, dtype=dt)
        opt = eng.make_options(nbNonzeroCoefs, toleranceResidualScale, toleranceSnr, nbBlocks, minCoefficients,
                               use_weights=weights is not None, coef_mode=self.coef_mode, method=self._method)
        res = eng.encode_host(np.ascontiguousarray(x, dtype=dt), opt, n_chunks=max(1, min(8, x.shape[0] // 32)))
        self.last_result = res
        codes = [res.to_csc(s, minCoefficients) for s in range(res.S)]
        residual = res.residual.numpy().astype(sequences.dtype if sequences.dtype.kind == 'f' else np.float64)
        if sequences.ndim == 2:
            residual = residual[:, :, 0]
        return codes, residual


class LoCOMP(ConvolutionalMatchingPursuit):
    """Drop-in for hsc.modeling.LoCOMP (:1191-1425): MP plus a local least-squares refit of the selected atom and
    the already-selected atoms sharing its support, on the device (csrc/locomp.cuh)."""
    _method = 1

    def __init__(self, verbose=False, device=None, coef_mode=1):
        super(LoCOMP, self).__init__(verbose, device, coef_mode)


class HierarchicalConvolutionalMatchingPursuit(SparseApproximator):
    """Drop-in for hsc.modeling.HierarchicalConvolutionalMatchingPursuit (:1427-1654): the level loop
    (each level's input is the previous level's dense code map, :1489), singleton weights
    (:1448-1450), redistribution (:1556-1594) and the input-level residual (:1596-1611)."""

    def __init__(self, method='cmp', device=None, coef_mode=1):
        self.method = method
        self.device = device
        self.coef_mode = coef_mode

    def _level_approximator(self):
        if self.method == 'cmp':
            return ConvolutionalMatchingPursuit(device=self.device, coef_mode=self.coef_mode)
        if self.method == 'locomp':
            return LoCOMP(device=self.device, coef_mode=self.coef_mode)
        raise Exception('Unsupported sparse coding method: %s' % (self.method))

    def _forward(self, sequence, coefficients, multilevelDict, toleranceSnr, nbBlocks, singletonWeight):
        input = sequence if len(coefficients) == 0 else np.asarray(coefficients[-1].todense())
        for level in range(len(coefficients), multilevelDict.getNbLevels()):
            if toleranceSnr is not None and isinstance(toleranceSnr, collections.abc.Iterable):
                targetSnr = toleranceSnr[level]
            else:
                targetSnr = toleranceSnr
            D = multilevelDict.getRawDictionary(level)
            nbSingletons = D.shape[0] - multilevelDict.countsNoSingletons[level]
            weights = np.ones((D.shape[0],), dtype=D.dtype)
            weights[:nbSingletons] = singletonWeight
            cmp = self._level_approximator()
            levelCoefficients, _ = ConvolutionalSparseCoder(D, cmp).encode(input, toleranceSnr=targetSnr, nbBlocks=nbBlocks,
                                                                           weights=weights)
            input = np.asarray(levelCoefficients.todense())
            coefficients.append(levelCoefficients)
        return coefficients

    def convertToDistributedCoefficients(self, coefficients):
        last = scipy.sparse.csc_matrix(coefficients[-1]).copy()
        out = []
        for level in range(len(coefficients)):
            if level < len(coefficients) - 1:
                nf = coefficients[level].shape[1]
                lvl = last[:, :nf]
                last = scipy.sparse.hstack((scipy.sparse.csc_matrix((last.shape[0], nf), dtype=last.dtype), last[:, nf:])).tocsc()
                lvl.eliminate_zeros()
            else:
                lvl = last
            out.append(lvl)
        assert len(out) == len(coefficients)
        assert np.sum([c.nnz for c in out]) == coefficients[-1].nnz
        return out

    def _calculateResidual(self, sequence, coefficients, multilevelDict):
        baseDict = multilevelDict.getBaseDictionary()
        shape = (coefficients[0].shape[0],) if baseDict.ndim == 2 else (coefficients[0].shape[0], baseDict.shape[-1])
        reconstruction = np.zeros(shape, dtype=coefficients[0].dtype)
        representations = multilevelDict.getMultiscaleDictionaries()
        for level in range(multilevelDict.getNbLevels()):
            reconstruction += reconstructSignal(coefficients[level], representations[level], device=self.device)
        return sequence - reconstruction

    def _postprocessCoefficients(self, coefficients, multilevelDict, returnDistributed=True):
        if returnDistributed:
            return self.convertToDistributedCoefficients(coefficients)
        out = []
        for level in range(multilevelDict.getNbLevels()):
            c = coefficients[level]
            if level < multilevelDict.getNbLevels() - 1:
                c = scipy.sparse.csc_matrix(c.shape, dtype=c.dtype)
            out.append(c)
        return out

    def computeCoefficients(self, sequence, multilevelDict, nbNonzeroCoefs=None, toleranceResidualScale=None,
                            toleranceSnr=None, nbBlocks=1, minCoefficients=None, singletonWeight=0.5,
                            returnDistributed=True, stopCondition=None):
        assert _is_multilevel_dictionary(multilevelDict)
        coefficients = self._forward(np.asarray(sequence), [], multilevelDict, toleranceSnr, nbBlocks, singletonWeight)
        coefficients = self._postprocessCoefficients(coefficients, multilevelDict, returnDistributed)
        residual = self._calculateResidual(np.asarray(sequence), coefficients, multilevelDict)
        return coefficients, residual

    def computeCoefficientsFromLevel(self, sequence, coefficients, multilevelDict, nbNonzeroCoefs=None,
                                     toleranceResidualScale=None, toleranceSnr=None, nbBlocks=1, minCoefficients=None,
                                     singletonWeight=0.5, stopCondition=None, returnDistributed=True):
        assert _is_multilevel_dictionary(multilevelDict)
        coefficients = copy.deepcopy(coefficients)
        coefficients = self._forward(np.asarray(sequence), coefficients, multilevelDict, toleranceSnr, nbBlocks, singletonWeight)
        return self._postprocessCoefficients(coefficients, multilevelDict, returnDistributed)


class ConvolutionalSparseCoder(object):
    """hsc/modeling.py:1656-1669."""

    def __init__(self, D, approximator):
        assert D.ndim == 2 or D.ndim == 3
        self.D = D
        self.approximator = approximator

    def encode(self, X, *args, **kwargs):
        assert X.ndim == 1 or X.ndim == 2
        return self.approximator.computeCoefficients(X, self.D, *args, **kwargs)

    def encodeBatch(self, X, *args, **kwargs):
        assert X.ndim == 2 or X.ndim == 3
        return self.approximator.computeCoefficientsBatch(X, self.D, *args, **kwargs)

    def reconstruct(self, coefficients):
        assert coefficients.ndim == 1 or coefficients.ndim == 2
        return reconstructSignal(coefficients, self.D, device=getattr(self.approximator, 'device', None))


class HierarchicalConvolutionalSparseCoder(object):
    """hsc/modeling.py:1671-1705."""

    def __init__(self, multilevelDict, approximator):
        assert _is_multilevel_dictionary(multilevelDict)
        if not multilevelDict.hasSingletonBases:
            multilevelDict = multilevelDict.withSingletonBases()
        self.multilevelDict = multilevelDict
        self.approximator = approximator

    def encode(self, sequence, *args, **kwargs):
        assert sequence.ndim == 1 or sequence.ndim == 2
        return self.approximator.computeCoefficients(sequence, self.multilevelDict, *args, **kwargs)

    def encodeFromLevel(self, sequence, coefficients, *args, **kwargs):
        assert len(coefficients) > 0
        return self.approximator.computeCoefficientsFromLevel(sequence, coefficients, self.multilevelDict, *args, **kwargs)

    def reconstruct(self, coefficients):
        assert len(coefficients) > 0
        baseDict = self.multilevelDict.getBaseDictionary()
        shape = (coefficients[0].shape[0],) if baseDict.ndim == 2 else (coefficients[0].shape[0], baseDict.shape[-1])
        signal = np.zeros(shape, dtype=coefficients[0].dtype)
        representations = self.multilevelDict.getMultiscaleDictionaries()
        for level in range(self.multilevelDict.getNbLevels()):
            signal += reconstructSignal(coefficients[level], representations[level],
                                        device=getattr(self.approximator, 'device', None))
        return signal
