"""Shared helpers of the parity harness (test infrastructure)."""
import ast
import os

import numpy as np
import scipy.sparse

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def case_kwargs(z, name):
    kw = ast.literal_eval(str(z[name + '_kw']))
    if 'weights' in kw and kw['weights'] is not None:
        kw['weights'] = np.asarray(kw['weights'], dtype=z[name + '_D'].dtype)
    return kw


def coo_sorted(m):
    c = scipy.sparse.coo_matrix(m)
    c.sum_duplicates()
    order = np.lexsort((c.col, c.row))
    return c.row[order].astype(np.int64), c.col[order].astype(np.int64), c.data[order].astype(np.float64)


def snr_db(x, residual):
    e = float(np.sum(np.square(np.asarray(residual, dtype=np.float64))))
    s = float(np.sum(np.square(np.asarray(x, dtype=np.float64))))
    return np.inf if e == 0.0 else 10.0 * np.log10(s / e)


class TraceComparison(object):
    """Step-wise comparison of two selection traces under the north-star rule: identical (t,k)
    sequence except at documented near-ties (relative correlation gap < tie_tol); after the first
    permitted flip the ORDER may differ, so the accumulated codes are compared instead."""

    def __init__(self, ref_t, ref_k, ref_c, got_t, got_k, got_c):
        self.ref = (np.asarray(ref_t), np.asarray(ref_k), np.asarray(ref_c, dtype=np.float64))
        self.got = (np.asarray(got_t), np.asarray(got_k), np.asarray(got_c, dtype=np.float64))
        n = min(len(self.ref[0]), len(self.got[0]))
        same = (self.ref[0][:n] == self.got[0][:n]) & (self.ref[1][:n] == self.got[1][:n])
        self.common_prefix = int(n if same.all() else np.argmin(same))
        self.n_ref = len(self.ref[0])
        self.n_got = len(self.got[0])

    @property
    def identical_sequence(self):
        return self.common_prefix == self.n_ref == self.n_got

    def prefix_coef_rel_err(self):
        n = self.common_prefix
        if n == 0:
            return 0.0
        r, g = self.ref[2][:n], self.got[2][:n]
        scale = np.maximum(np.abs(r), 1e-3 * np.max(np.abs(r)))
        return float(np.max(np.abs(r - g) / scale))

    def divergence_gap(self):
        """Relative gap between the two candidates at the first divergent step, measured on the
        reference's coefficients: |c_ref(step)| vs the magnitude the other side picked."""
        n = self.common_prefix
        if n >= min(self.n_ref, self.n_got):
            return 0.0
        a, b = abs(self.ref[2][n]), abs(self.got[2][n])
        return float(abs(a - b) / max(a, b, 1e-300))


def accumulate(t, k, c, shape):
    m = scipy.sparse.coo_matrix((np.asarray(c, dtype=np.float64), (np.asarray(t), np.asarray(k))), shape=shape)
    m = m.tocsc()
    return m


def code_diff(ref, got, rel=1e-5):
    """|delta| <= rel * max(|c_ref|, scale) on the union support.  Returns
    (max_violation_ratio, n_support_mismatch)."""
    ref = scipy.sparse.csc_matrix(ref)
    got = scipy.sparse.csc_matrix(got)
    d = (ref - got).tocoo()
    if ref.nnz == 0:
        return (0.0 if got.nnz == 0 else np.inf), got.nnz
    scale = float(np.max(np.abs(ref.data)))
    refd = np.abs(np.asarray(ref[d.row, d.col])).ravel() if d.nnz else np.zeros(0)
    tol = rel * np.maximum(refd, 1e-2 * scale)
    ratio = float(np.max(np.abs(d.data) / tol)) if d.nnz else 0.0
    sr = set(zip(*ref.nonzero()))
    sg = set(zip(*got.nonzero()))
    return ratio, len(sr ^ sg)
