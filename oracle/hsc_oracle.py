"""CPU oracle for the convolutional matching-pursuit path of sbrodeur/hierarchical-sparse-coding.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module, and only as the
checker (or as the timed CPU arm).  The product package never imports it and has no CPU fallback.

This is a NumPy restatement of the reference algorithm (all citations relative to the reference
tree, `hsc/...`).  It is written from the behaviour, not from the text, of the reference: a single
centre/clip helper replaces the reference's three copies of the even/odd boundary code, events are
kept in a plain dict instead of a `scipy.sparse.lil_matrix` (optional `bookkeeping='lil'`
reproduces the reference's cost profile for the CPU-baseline timing), and every run returns the
step-wise trace `(t, k, c)` that the parity harness compares against the CUDA engine.

PARITY PINNED: `tests/test_oracle_golden.py` checks this file against (a) the golden vectors of
the reference's own tests (tests/hsc/test_utils.py:113-218, tests/hsc/test_modeling.py:272-325,
:379-396, :712-725, :774-821), re-stated under `tests/golden/`, (b) traces recorded from the
unmodified reference run in the dev container (`tests/golden/make_golden.py`), and (c) - when
`/root/reference` is mounted - the live reference on seeded random inputs, bit for bit.

The dense products go through `np.dot` on operands of exactly the shapes / memory order the
reference builds (hsc/modeling.py:181-187), so on the same NumPy/BLAS the numbers are
bit-identical to the reference's, not merely close.
"""
import collections.abc
import math

import numpy as np
import scipy.linalg
import scipy.sparse
from numpy.lib.stride_tricks import sliding_window_view


# ----------------------------------------------------------------------------------------------
# Index convention (hsc/utils.py:76-161, hsc/modeling.py:845-858)
# ----------------------------------------------------------------------------------------------

def centre_offset(width):
    """Index of the 'centre' tap inside a filter of `width` taps: width/2-1 (even), width//2 (odd).
    hsc/utils.py:83-99 (peek), hsc/modeling.py:846-852 (Atom.getPositionSpanIndices)."""
    return width // 2 - 1 if width % 2 == 0 else width // 2


def clipped_span(width, t, length):
    """For an element of `width` taps centred at `t` on a signal of `length` samples, returns
    (lo, hi, elo, ehi): signal[lo:hi] is covered by element[elo:ehi]; hi <= lo means 'nothing'.
    One statement of the boundary rule that hsc/utils.py repeats in peek (:76-101),
    overlapAdd (:103-131) and overlapReplace (:133-161)."""
    start = t - centre_offset(width)
    lo = max(start, 0)
    hi = min(start + width, length)
    return lo, hi, lo - start, hi - start


def peek(signal, width, t):
    """hsc/utils.py:76-101 -- clipped read of the window centred at t."""
    lo, hi, _, _ = clipped_span(width, t, signal.shape[0])
    if hi <= lo:
        return np.array([], dtype=signal.dtype)
    return signal[lo:hi]


def overlap_add(signal, element, t):
    """hsc/utils.py:103-131 -- in-place clipped add of `element` centred at t."""
    lo, hi, elo, ehi = clipped_span(element.shape[0], t, signal.shape[0])
    if hi > lo:
        signal[lo:hi] += element[elo:ehi]
    return signal


def overlap_replace(signal, element, t):
    """hsc/utils.py:133-161 -- in-place clipped overwrite by `element` centred at t."""
    lo, hi, elo, ehi = clipped_span(element.shape[0], t, signal.shape[0])
    if hi > lo:
        signal[lo:hi] = element[elo:ehi]
    return signal


def normalize(X, axis=None):
    """hsc/utils.py:67-74 -- unit L2 norm per leading-axis item, zero-norm safe."""
    if axis is None and X.ndim > 1:
        axis = tuple(range(1, X.ndim))
    n = np.sqrt(np.sum(np.square(X), axis=axis, keepdims=True))
    n = np.where(n > 0.0, n, np.ones_like(n))
    return X / n


# ----------------------------------------------------------------------------------------------
# Correlation (hsc/modeling.py:149-188)
# ----------------------------------------------------------------------------------------------

def _as_tf(sequence):
    sequence = np.asarray(sequence)
    return sequence.reshape(sequence.shape[0], -1)


def correlate(sequence, filters, padding='valid'):
    """Multichannel 1-D cross-correlation (no flip) of [T,F] with [K,L,F] -> [T',K].
    'same' zero-pads (L/2-1, L/2) for even L and (L//2, L//2) for odd L (hsc/modeling.py:157-164).
    The product is one np.dot of the [T', F*L] window matrix (feature-major, tap-minor, as the
    reference's as_strided view orders it, :181-186) with filters.T reshaped to [F*L, K] (:187)."""
    x = _as_tf(sequence)
    L = filters.shape[1]
    if padding == 'same':
        before = centre_offset(L)
        x = np.pad(x, [(before, L - 1 - before), (0, 0)], mode='constant')
    elif padding != 'valid':
        raise Exception('Padding not supported: %s' % (padding))
    F = 1 if filters.ndim == 2 else filters.shape[-1]
    assert F == x.shape[-1]
    windows = sliding_window_view(x, L, axis=0)            # [T', F, L], strides as in :181-182
    nq = int(np.prod(filters.shape[1:]))
    A = windows.reshape((windows.shape[0], nq))            # materialised im2col copy (:186)
    Bm = filters.T.reshape(nq, filters.shape[0])           # [F*L, K] (:187)
    return np.dot(A, Bm)


def reconstruct(coefficients, D):
    """Sparse decoder: x_hat = sum_n c_n * D[k_n] centred at t_n (hsc/modeling.py:226-245).
    Dense input goes through the same scatter (the reference uses an FFT convolution there,
    :247-258, equal up to rounding; reference test tests/hsc/test_modeling.py:774-821)."""
    D3 = D[:, :, None] if D.ndim == 2 else D
    cx = scipy.sparse.coo_matrix(coefficients)
    out = np.zeros((cx.shape[0], D3.shape[-1]), dtype=cx.dtype)
    for t, k, c in zip(cx.row, cx.col, cx.data):
        if c != 0.0:
            overlap_add(out, c * D3[k], t)
    return out[:, 0] if D.ndim == 2 else out


# ----------------------------------------------------------------------------------------------
# Atom selection (hsc/modeling.py:899-982)
# ----------------------------------------------------------------------------------------------

def block_geometry(T, L, nb_blocks, offset):
    """Block size / count / front padding of the block-wise selection (hsc/modeling.py:908-928).
    Returns (blockSize, nBlocks, padFront)."""
    if nb_blocks == 'auto':
        bs = 4 * L
    else:
        bs = int(np.floor(T / float(nb_blocks)))
    if bs % 2 == 1:
        bs += 1
    nb = int(np.ceil(T / float(bs)))
    if offset:
        return bs, nb + 1, bs // 2
    return bs, nb, 0


def select_atoms(inner, L, nb_blocks=1, offset=False, null_thres=0.0, weights=None):
    """Returns the list [(t, k, c)] of the pass, in application order.
    nb_blocks == 1: global argmax of |inner*w| in row-major order (lowest t, then lowest k wins a
    tie), coefficient read from the UNWEIGHTED map, dropped if |c| <= null_thres (:965-975).
    Otherwise one argmax per time block (blocks shifted by half a block on 'offset' passes),
    range filter, null filter, interference filter on the time gaps, sort by |c| descending via
    argsort()[::-1] (:908-963)."""
    scores = inner if weights is None else inner * np.asarray(weights)[None, :]
    T, K = inner.shape
    if nb_blocks == 'auto' or nb_blocks > 1:
        bs, nb, front = block_geometry(T, L, nb_blocks, offset)
        back = nb * bs - front - T
        padded = np.pad(scores, [(front, back), (0, 0)], mode='constant')
        flat = np.abs(padded.reshape(nb, bs * K))
        best = np.argmax(flat, axis=1)
        t = best // K + np.arange(nb) * bs - front
        k = best % K
        keep = np.where((t >= 0) & (t <= T - 1))[0]
        t, k = t[keep], k[keep]
        c = inner[t, k]
        keep = np.where(np.abs(c) > null_thres)
        t, k, c = t[keep], k[keep], c[keep]
        gaps = t[1:] - t[:-1]
        far = np.where(gaps >= L)[0]
        if len(far) > 0:
            keep = np.concatenate(([0], far + 1))
            t, k, c = t[keep], k[keep], c[keep]
        order = np.argsort(np.abs(c))[::-1]
        t, k, c = t[order], k[order], c[order]
    else:
        best = int(np.argmax(np.abs(scores)))
        t = np.array([best // K])
        k = np.array([best % K])
        c = inner[t, k]
        keep = np.where(np.abs(c) > null_thres)
        t, k, c = t[keep], k[keep], c[keep]
    return [(int(a), int(b), cc) for a, b, cc in zip(t, k, c)]


# ----------------------------------------------------------------------------------------------
# Local updates (hsc/modeling.py:996-1051)
# ----------------------------------------------------------------------------------------------

def subtract_atom(residual, t, k, c, D3):
    """r[span] -= c*D[k]; returns (E_before - E_after) over the clipped span
    (hsc/modeling.py:1002-1005)."""
    L = D3.shape[1]
    before = np.sum(np.square(peek(residual, L, t)))
    overlap_add(residual, -c * D3[k], t)
    after = np.sum(np.square(peek(residual, L, t)))
    return before - after


def refresh_window(inner, residual, t, D3):
    """Re-correlates the 3L-2 residual samples around t -- REFLECT-padded where they overhang the
    signal (np.pad mode='reflect', hsc/modeling.py:1046), unlike the zero-padded initial map --
    and overwrites map rows [t-(L-1), t+(L-1)] (clipped) (:1018-1051)."""
    L = D3.shape[1]
    T = residual.shape[0]
    first = t - centre_offset(L) - (L - 1)
    last = t + L // 2 + (L - 1)
    lo, hi = max(first, 0), min(last, T - 1)
    padded = np.pad(residual[lo:hi + 1], [(lo - first, last - hi), (0, 0)], mode='reflect')
    local = correlate(padded, D3, 'valid')
    assert local.shape[0] == 2 * L - 1
    overlap_replace(inner, local, t)


# ----------------------------------------------------------------------------------------------
# Sparse bookkeeping: dict (fast, default) or scipy LIL (the reference's cost profile)
# ----------------------------------------------------------------------------------------------

class _DictCode(object):
    def __init__(self, T, K):
        self.shape = (T, K)
        self.d = {}

    def get(self, t, k):
        return self.d.get((t, k), 0.0)

    def add(self, t, k, c):
        self.d[(t, k)] = self.d.get((t, k), 0.0) + float(c)

    def nnz(self):
        return len(self.d)

    def items(self):
        return [(t, k, v) for (t, k), v in self.d.items()]

    def window_items(self, lo, hi):
        return [(t, k, v) for (t, k), v in self.d.items() if lo <= t <= hi]

    def to_coo(self, min_coef):
        it = self.items()
        if min_coef is not None:
            it = [e for e in it if abs(e[2]) >= min_coef]
        it = [e for e in it if e[2] != 0.0]
        if len(it) == 0:
            return scipy.sparse.csc_matrix(self.shape, dtype=np.float64)
        t, k, v = zip(*it)
        return scipy.sparse.coo_matrix((np.array(v, dtype=np.float64), (np.array(t), np.array(k))),
                                       shape=self.shape).tocsc()


class _LilCode(object):
    """Same operations on a scipy.sparse.lil_matrix, with the reference's access pattern
    (fancy-index `+=`, hsc/modeling.py:984-994, :1106) so that the timed CPU baseline pays what
    the reference pays."""

    def __init__(self, T, K):
        self.shape = (T, K)
        self.m = scipy.sparse.lil_matrix((T, K))

    def get(self, t, k):
        return self.m[t, k]

    def add(self, t, k, c):
        self.m[[t], [k]] += np.array([c], dtype=self.m.dtype)

    def nnz(self):
        return self.m.nnz

    def items(self):
        cx = self.m.tocoo()
        return list(zip(cx.row.tolist(), cx.col.tolist(), cx.data.tolist()))

    def window_items(self, lo, hi):
        cx = self.m[lo:hi + 1, :].tocoo()
        return [(lo + int(r), int(c), v) for r, c, v in zip(cx.row, cx.col, cx.data)]

    def to_coo(self, min_coef):
        m = self.m
        if min_coef is not None:
            clipped = scipy.sparse.lil_matrix(self.shape)
            cx = m.tocoo()
            keep = np.where(np.abs(cx.data) >= min_coef)
            clipped[cx.row[keep], cx.col[keep]] = cx.data[keep]
            m = clipped
        m = m.tocsc()
        m.eliminate_zeros()
        return m


class Trace(object):
    """What the parity harness compares: the ordered selections and the stop state."""

    def __init__(self):
        self.events = []          # (t, k, c) in application order (LoCOMP: every refitted group atom)
        self.selections = 0
        self.passes = 0
        self.nnz = 0
        self.duplicates = 0
        self.energy_signal = None
        self.energy_residual = None
        self.stop = None
        self.snapshot = None      # mp_encode(snapshot_event=n): state at the start of the pass that produces event n

    def arrays(self):
        if len(self.events) == 0:
            return (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float64))
        t, k, c = zip(*self.events)
        return np.array(t, np.int64), np.array(k, np.int64), np.array(c, np.float64)


def _prepare(sequence, D):
    assert sequence.ndim == 1 or sequence.ndim == 2
    assert D.ndim == 2 or D.ndim == 3
    squeeze = sequence.ndim == 1 or D.ndim == 2
    x = sequence[:, None] if sequence.ndim == 1 else sequence
    D3 = D[:, :, None] if D.ndim == 2 else D
    return np.asarray(x), np.asarray(D3), squeeze


def weak_atom_filter(atoms, residual, L, energy_signal, tol_snr, n_samples):
    """Drops atoms whose local mean energy is already below the mean residual energy of the SNR
    target (only when a pass holds more than one atom) (hsc/modeling.py:1090-1099)."""
    target = energy_signal / (10.0 ** (tol_snr / 10.0)) / n_samples
    return [a for a in atoms if np.mean(np.square(peek(residual, L, a[0]))) >= target]


# ----------------------------------------------------------------------------------------------
# Matching pursuit (hsc/modeling.py:1053-1186)
# ----------------------------------------------------------------------------------------------

def mp_encode(sequence, D, nbNonzeroCoefs=None, toleranceResidualScale=None, toleranceSnr=None,
              nbBlocks=1, minCoefficients=1e-16, weights=None, stopCondition=None,
              bookkeeping='dict', max_events=None, return_trace=False, snapshot_event=None):
    """ConvolutionalMatchingPursuit.computeCoefficients restated.  Returns
    (csc_matrix[T,K] float64, residual like the input) and, with return_trace, the Trace.
    `max_events` (not in the reference) bounds the number of applied atoms for bounded timing
    samples; it acts like an extra stop tested after each atom.  `snapshot_event=n` (parity harness)
    stops at the START of the selection pass that produces event number n (0-based) and leaves the
    correlation map, the residual and that pass's atom list in Trace.snapshot: the state the
    reference ranks its candidates on at that step."""
    x, D3, squeeze = _prepare(sequence, D)
    eps = np.finfo(D3.dtype).eps
    T, K, L = x.shape[0], D3.shape[0], D3.shape[1]

    energy_signal = np.sum(np.square(x))
    residual = np.copy(x)
    energy = energy_signal
    code = (_LilCode if bookkeeping == 'lil' else _DictCode)(T, K)
    inner = correlate(residual, D3, 'same')

    tr = Trace()
    tr.energy_signal = energy_signal
    offset = False
    done = False
    while not done:
        atoms = select_atoms(inner, L, nbBlocks, offset, minCoefficients, weights)
        if toleranceSnr is not None and len(atoms) > 1:
            atoms = weak_atom_filter(atoms, residual, L, energy_signal, toleranceSnr, x.size)
        if snapshot_event is not None and (len(tr.events) + len(atoms) > snapshot_event or len(atoms) == 0):
            tr.snapshot = dict(inner=inner, residual=residual, first_event=len(tr.events), atoms=atoms, offset=offset)
            tr.stop = 'snapshot'
            break
        for (t, k, c) in atoms:
            if np.abs(code.get(t, k)) > 0.0:
                tr.duplicates += 1
            elif np.abs(c) > 0.0:
                tr.nnz += 1
            code.add(t, k, c)
            energy = energy - (0.0 + subtract_atom(residual, t, k, c, D3))
            refresh_window(inner, residual, t, D3)
            tr.events.append((t, k, float(c)))
            if energy < eps:
                done, tr.stop = True, 'energy'
                break
            snr = 10.0 * np.log10(energy_signal / energy)
            if nbNonzeroCoefs is not None and tr.nnz >= nbNonzeroCoefs:
                done, tr.stop = True, 'nnz'
                break
            if toleranceSnr is not None and snr >= toleranceSnr:
                done, tr.stop = True, 'snr'
                break
            if max_events is not None and len(tr.events) >= max_events:
                done, tr.stop = True, 'max_events'
                break
        scale = np.max(np.abs(residual))                     # full scan every pass (:1145)
        if toleranceResidualScale is not None and scale <= toleranceResidualScale:
            done = True
            tr.stop = tr.stop or 'scale'
        if len(atoms) == 0:
            done = True
            tr.stop = tr.stop or 'empty'
        if stopCondition is not None and stopCondition(x, residual, code.to_coo(None)):
            done = True
            tr.stop = tr.stop or 'callback'
        tr.passes += 1
        offset = not offset

    tr.energy_residual = energy
    out = code.to_coo(minCoefficients)
    res = residual[:, 0] if squeeze else residual
    if return_trace:
        return out, res, tr
    return out, res


# ----------------------------------------------------------------------------------------------
# LoCOMP (hsc/modeling.py:1191-1425)
# ----------------------------------------------------------------------------------------------

def atom_span(L, t, T=None):
    """Atom.getPositionSpanIndices (hsc/modeling.py:845-858): inclusive [start, end]."""
    s = t - centre_offset(L)
    e = t + L // 2
    if T is not None:
        s, e = max(s, 0), min(e, T - 1)
    return s, e


def common_support(code, t, k, L, T):
    """_findCommonSupportAtoms (hsc/modeling.py:1221-1239).  The reference compares the
    slice-RELATIVE row with the absolute atom position and joins the two tests with `and`
    (:1238): an existing entry is kept iff (row - lo) != t and column != k.  Replicated as is."""
    s, e = atom_span(L, t, T)
    lo = max(s - L // 2, 0)
    hi = min(e + (L // 2 - 1 if L % 2 == 0 else L // 2), T)
    items = sorted(code.window_items(lo, min(hi, T - 1)))     # coo of a LIL slice is row-major
    return [(tt, kk, vv) for (tt, kk, vv) in items if (tt - lo) != t and kk != k]


def support_dictionary(residual, atoms, D3):
    """_getDictionaryFromSupportAtoms (hsc/modeling.py:1241-1261)."""
    L = D3.shape[1]
    T = residual.shape[0]
    lo = min(atom_span(L, a[0], T)[0] for a in atoms)
    hi = 0
    for a in atoms:
        hi = max(hi, atom_span(L, a[0], T)[1])
    n = hi - lo + 1
    rows = []
    for a in atoms:
        buf = np.zeros((n,) + residual.shape[1:], dtype=D3.dtype)
        overlap_add(buf, D3[a[1]], a[0] - lo)
        rows.append(buf)
    return np.stack(rows), residual[lo:hi + 1]


def locomp_encode(sequence, D, nbNonzeroCoefs=None, toleranceResidualScale=None, toleranceSnr=None,
                  nbBlocks=1, minCoefficients=1e-16, weights=None, stopCondition=None,
                  bookkeeping='dict', max_events=None, return_trace=False):
    """LoCOMP.computeCoefficients restated (hsc/modeling.py:1263-1425): MP plus, per selected
    atom, a least-squares refit (pinv) of the atom and the already-selected atoms sharing its
    support; the fitted values are ADDED (they are fitted to the residual) (:1322-1353)."""
    x, D3, squeeze = _prepare(sequence, D)
    eps = np.finfo(D3.dtype).eps
    T, K, L = x.shape[0], D3.shape[0], D3.shape[1]

    energy_signal = np.sum(np.square(x))
    residual = np.copy(x)
    energy = energy_signal
    code = (_LilCode if bookkeeping == 'lil' else _DictCode)(T, K)
    inner = correlate(residual, D3, 'same')

    tr = Trace()
    tr.energy_signal = energy_signal
    offset = False
    done = False
    while not done:
        atoms = select_atoms(inner, L, nbBlocks, offset, minCoefficients, weights)
        if toleranceSnr is not None and len(atoms) > 1:
            atoms = weak_atom_filter(atoms, residual, L, energy_signal, toleranceSnr, x.size)
        for (t, k, c) in atoms:
            last_energy = energy
            group = [(t, k, c)]
            others = common_support(code, t, k, L, T)
            if len(others) > 0:
                group = group + others
                Dsup, rsup = support_dictionary(residual, group, D3)
                flat = Dsup.reshape((Dsup.shape[0], -1))
                fit = np.dot(np.linalg.pinv(flat).T, rsup.flatten()).flatten()
                group = [(g[0], g[1], f) for g, f in zip(group, fit)]
            # the reference adds all group coefficients in ONE fancy-index `+=` (:1338), which for
            # repeated (t,k) pairs keeps only the last; groups never repeat a pair (dict keys).
            for (tt, kk, cc) in group:
                code.add(tt, kk, cc)
            loss = 0.0
            for (tt, kk, cc) in group:
                loss += subtract_atom(residual, tt, kk, cc, D3)
            energy = energy - loss
            for (tt, kk, cc) in group:
                refresh_window(inner, residual, tt, D3)
            tr.events.extend((tt, kk, float(cc)) for (tt, kk, cc) in group)
            tr.selections += 1
            if energy < eps:
                done, tr.stop = True, 'energy'
                break
            snr = 10.0 * np.log10(energy_signal / energy)
            if nbNonzeroCoefs is not None and code.nnz() >= nbNonzeroCoefs:
                done, tr.stop = True, 'nnz'
                break
            if toleranceSnr is not None and snr >= toleranceSnr:
                done, tr.stop = True, 'snr'
                break
            if np.abs(last_energy - energy) < eps:
                done, tr.stop = True, 'stall'
                break
            if max_events is not None and tr.selections >= max_events:
                done, tr.stop = True, 'max_events'
                break
        scale = np.max(np.abs(residual))
        if toleranceResidualScale is not None and scale <= toleranceResidualScale:
            done = True
            tr.stop = tr.stop or 'scale'
        if len(atoms) == 0:
            done = True
            tr.stop = tr.stop or 'empty'
        if stopCondition is not None and stopCondition(code.to_coo(None)):
            done = True
            tr.stop = tr.stop or 'callback'
        tr.passes += 1
        offset = not offset

    tr.nnz = code.nnz()
    tr.energy_residual = energy
    out = code.to_coo(minCoefficients)
    res = residual[:, 0] if squeeze else residual
    if return_trace:
        return out, res, tr
    return out, res


# ----------------------------------------------------------------------------------------------
# Hierarchical MP (hsc/modeling.py:1427-1654) on plain arrays
# ----------------------------------------------------------------------------------------------

def distribute_levels(codes):
    """convertToDistributedCoefficients (hsc/modeling.py:1556-1594): columns [:K_l] of the LAST
    level's code are the pass-through (singleton) events of level l; they are cut out level by
    level so the total nnz is conserved (:1592)."""
    last = scipy.sparse.csc_matrix(codes[-1]).copy()
    out = []
    for level in range(len(codes)):
        if level < len(codes) - 1:
            nf = codes[level].shape[1]
            lvl = last[:, :nf]
            last = scipy.sparse.hstack((scipy.sparse.csc_matrix((last.shape[0], nf), dtype=last.dtype),
                                        last[:, nf:])).tocsc()
            lvl.eliminate_zeros()
        else:
            lvl = last
        out.append(lvl)
    assert sum(c.nnz for c in out) == codes[-1].nnz
    return out


def hierarchical_residual(sequence, codes, representations):
    """_calculateResidual (hsc/modeling.py:1596-1611): x - sum_l decode(code_l, input-level
    representations of level l)."""
    base = representations[0]
    shape = (codes[0].shape[0],) if base.ndim == 2 else (codes[0].shape[0], base.shape[-1])
    rec = np.zeros(shape, dtype=codes[0].dtype)
    for lvl in range(len(representations)):
        rec += reconstruct(codes[lvl], representations[lvl])
    return sequence - rec


def hierarchical_encode(sequence, raw_dictionaries, counts_no_singletons, representations,
                        toleranceSnr=None, nbBlocks=1, singletonWeight=0.5, returnDistributed=True,
                        method='cmp', from_codes=None, return_traces=False, bookkeeping='dict'):
    """HierarchicalConvolutionalMatchingPursuit.computeCoefficients on plain arrays
    (hsc/modeling.py:1432-1492 forward phase, :1613-1634 post-processing, :1636-1643).
    Level l encodes the DENSE float64 code map of level l-1 as an F=K_{l-1}-channel signal
    (:1489) with weights[:nbSingletons] = singletonWeight (:1448-1450).  `from_codes` resumes like
    computeCoefficientsFromLevel (:1645-1654)."""
    encode = mp_encode if method == 'cmp' else locomp_encode
    codes = [] if from_codes is None else [c.copy() for c in from_codes]
    traces = []
    inp = sequence if not codes else np.asarray(codes[-1].todense())
    for level in range(len(codes), len(raw_dictionaries)):
        if toleranceSnr is not None and isinstance(toleranceSnr, collections.abc.Iterable):
            target = toleranceSnr[level]
        else:
            target = toleranceSnr
        D = raw_dictionaries[level]
        n_single = D.shape[0] - int(counts_no_singletons[level])
        w = np.ones((D.shape[0],), dtype=D.dtype)
        w[:n_single] = singletonWeight
        c, _, tr = encode(inp, D, toleranceSnr=target, nbBlocks=nbBlocks, weights=w,
                          return_trace=True, bookkeeping=bookkeeping)
        traces.append(tr)
        inp = np.asarray(c.todense())
        codes.append(c)
    if returnDistributed:
        codes = distribute_levels(codes)
    else:
        codes = [scipy.sparse.csc_matrix(c.shape, dtype=c.dtype) if i < len(codes) - 1 else c
                 for i, c in enumerate(codes)]
    if from_codes is not None:
        return (codes, traces) if return_traces else codes
    res = hierarchical_residual(sequence, codes, representations)
    return (codes, res, traces) if return_traces else (codes, res)


# ----------------------------------------------------------------------------------------------
# K-SVD dictionary update consuming the MP codes (hsc/modeling.py:593-636)
# ----------------------------------------------------------------------------------------------

def pca_first_component(windows):
    """hsc/modeling.py:48-80 with k=1: more than one row -> the rows are mean-centred IN PLACE (:54, the caller's
    projection at :625 therefore uses the centred windows) and the eigenvector of the largest eigenvalue of their
    covariance is returned (:57-69; the sign is LAPACK's); one row -> that row, L2-normalised, not centred (:76-78)."""
    m = windows.shape[0]
    if m > 1:
        windows -= windows.mean(axis=0)
        R = np.cov(windows, rowvar=False)
        evals, evecs = scipy.linalg.eigh(np.atleast_2d(R))
        return evecs[:, np.argsort(evals)[::-1][0]]
    return normalize(windows)[0]


def ksvd_dictionary_update(coefficients, D, use_pca=False):
    """One dictionary-update stage (hsc/modeling.py:594-633), Gauss-Seidel over the filters:
    zero column k of the code, decode WITHOUT filter k (the reference decodes, it does not form
    x minus the decode: author's TODO at :606), gather the length-L windows centred at the
    column's atoms, rank-1 SVD -> new filter (first left singular vector) and coefficients
    (s0 * first right singular vector).  The sign of the pair is LAPACK's choice.  use_pca (:618-625): first
    principal component of the mean-centred windows instead, coefficients = centred windows . component.
    Returns (D_new, coefficients_new (lil), alpha = ||D_new - D||_F)."""
    D = np.array(D, copy=True)
    old = np.copy(D)
    coefficients = scipy.sparse.lil_matrix(coefficients)
    L = D.shape[1]
    half = L // 2
    for k in range(D.shape[0]):
        idx = coefficients[:, [k]].nonzero()[0]
        if len(idx) == 0:
            continue
        coefficients[idx, k * np.ones_like(idx)] = 0.0
        err = reconstruct(coefficients.tocsc(), D)
        padded = np.pad(err, [(half, half)] + [(0, 0)] * (err.ndim - 1), mode='constant')
        starts = half + idx - centre_offset(L)
        win = np.stack([padded[s:s + L] for s in starts]).reshape(len(idx), -1)
        if use_pca:                                                                  # :618-625
            win = np.ascontiguousarray(win, dtype=np.float64)
            evec = pca_first_component(win)
            D[k, :] = evec.reshape(D.shape[1:])
            coefficients[idx, k * np.ones_like(idx)] = np.dot(win, evec)
            continue
        U, s, Vh = scipy.linalg.svd(win.T, full_matrices=False)
        D[k, :] = U[:, 0].reshape(D.shape[1:])
        coefficients[idx, k * np.ones_like(idx)] = Vh.T[:, 0] * s[0]
    alpha = math.sqrt(np.sum(np.square(D - old)))
    return D, coefficients, alpha


# ----------------------------------------------------------------------------------------------
# Convolutional k-means learner (hsc/modeling.py:420-526)
# ----------------------------------------------------------------------------------------------

def kmeans_assign(windows, D):
    """Assignment step (:455-470): correlate every training window with the centroids at its 'valid'
    positions, first maximum of |similarity| over the flattened [position][centroid] scores.
    Returns (positions[B] = first sample of the best patch, assignments[B], patches[B,W(,F)])."""
    W = D.shape[1]
    w3 = windows[:, :, None] if windows.ndim == 2 else windows
    D3 = D[:, :, None] if D.ndim == 2 else D
    ip = np.stack([correlate(w, D3, 'valid') for w in w3])                       # [B, Tw-W+1, K]
    flat = np.argmax(np.abs(ip.reshape(ip.shape[0], -1)), axis=1)
    pos, idx = np.unravel_index(flat, ip.shape[1:])
    patches = np.stack([windows[b, pos[b]:pos[b] + W] for b in range(windows.shape[0])])
    return pos, idx, patches


def kmeans_iteration(windows, D, resetMethod='noise', nbAveragedPatches=8):
    """One iteration of _train_kmean (:455-517).  A centroid counts as EMPTY when
    `np.any(np.where(assignments == c))` is False (:479-481): the test is on the INDICES of its
    windows, so a centroid whose only window is window 0 is reset too.  Resets consume np.random in
    centroid order (:485-491).  Returns (newD, alpha, nbResets, positions, assignments)."""
    pos, idx, patches = kmeans_assign(windows, D)
    n_resets = 0
    cents = []
    for c in range(D.shape[0]):
        assigned = np.where(idx == c)
        if np.any(assigned):
            centroid = np.mean(normalize(patches[assigned]), axis=0)
        else:
            if resetMethod == 'random_samples':
                centroid = patches[np.random.randint(low=0, high=patches.shape[0])]
            elif resetMethod == 'random_samples_average':
                centroid = np.mean(patches[np.random.randint(low=0, high=patches.shape[0], size=(nbAveragedPatches,))], axis=0)
            elif resetMethod == 'noise':
                centroid = np.random.uniform(low=-1.0, high=1.0, size=patches.shape[1:])
            else:
                raise Exception('Unsupported reset method: %s' % (resetMethod))
            n_resets += 1
        if np.sqrt(np.sum(np.square(centroid))) == 0.0:                         # :499-502
            centroid = centroid + 1e-9
        cents.append(centroid)
    newD = normalize(np.stack(cents))
    alpha = np.sqrt(np.sum(np.square(D - newD)))
    return newD, alpha, n_resets, pos, idx


def kmeans_train(data, k, windowSize, nbRandomWindows, maxIterations=100, tolerance=0.0, initMethod='random_samples',
                 resetMethod='noise', nbAveragedPatches=8):
    """_train_kmean end to end with the reference's np.random call sequence (:426-429: training windows of twice the
    centroid length, then the initial centroids)."""
    lo = np.random.randint(low=0, high=data.shape[0] - 2 * windowSize, size=(nbRandomWindows,))     # :88
    windows = np.stack([data[i:i + 2 * windowSize] for i in lo])
    if initMethod == 'noise':                                                                         # :320-321
        shape = (k, windowSize) + tuple(data.shape[1:])
        D = normalize(np.random.uniform(low=np.min(data), high=np.max(data), size=(k, windowSize, 1 if data.ndim == 1 else data.shape[-1])))
        D = D.reshape(shape)
    elif initMethod == 'random_samples':
        i0 = np.random.randint(low=0, high=data.shape[0] - windowSize, size=(k,))
        D = normalize(np.stack([data[i:i + windowSize] for i in i0]))
    else:
        raise Exception('Unsupported initialization method: %s' % (initMethod))
    n = 0
    alpha = tolerance + 1.0
    history = []
    while n < maxIterations and alpha > tolerance:
        D, alpha, resets, _, _ = kmeans_iteration(windows, D, resetMethod, nbAveragedPatches)
        history.append((float(alpha), resets))
        n += 1
    return D, history


# ----------------------------------------------------------------------------------------------
# Event list <-> sparse code converters (hsc/dataset.py:798-824)
# ----------------------------------------------------------------------------------------------

EVENT_DTYPE = np.dtype('int32,int32,int32,float32')


def sparse_matrices_to_events(coefficients):
    """convertSparseMatricesToEvents (hsc/dataset.py:798-811): (t, level, index, coefficient) records of all levels,
    sorted by time with a STABLE sort (ties keep level order, then the COO order of the level)."""
    events = []
    for level, c in enumerate(coefficients):
        c = c.tocoo()
        events.extend(zip(c.row, level * np.ones_like(c.row), c.col, c.data))
    events = sorted(events, key=lambda e: e[0])
    return np.array(events, dtype=EVENT_DTYPE)


def events_to_sparse_matrices(events, counts, sequenceLength):
    """convertEventsToSparseMatrices (hsc/dataset.py:813-824): one csr_matrix [sequenceLength, counts[level]] per level."""
    t = np.array([e[0] for e in events], dtype=int)
    lv = np.array([e[1] for e in events], dtype=int)
    f = np.array([e[2] for e in events], dtype=int)
    v = np.array([e[3] for e in events], dtype=events.dtype[-1])
    out = []
    for level, count in enumerate(counts):
        m = np.where(lv == level)
        out.append(scipy.sparse.coo_matrix((v[m], (t[m], f[m])), shape=(sequenceLength, count)).tocsr())
    return out
