"""Data formats either side of the matching-pursuit path: the reference keeps codes either as one
sparse matrix [T, K_level] per level or as a flat, time-sorted list of events (t, level, filter, coefficient)
(hsc/dataset.py:798-824, dtype 'int32,int32,int32,float32').  The engine's own output is already an event list per
signal (EncodeResult: centre position, filter, coefficient in selection order), so both converters are vectorised
array shuffles on the host; nothing here touches the device.
"""
import numpy as np
import scipy.sparse

EVENT_DTYPE = np.dtype('int32,int32,int32,float32')


def convertSparseMatricesToEvents(coefficients):
    """hsc/dataset.py:798-811.  All levels' nonzeros as (t, level, index, coefficient) records sorted by time; the sort is
    stable, so simultaneous events keep level order and, inside a level, the COO order of its matrix."""
    ts, lv, ks, vs = [], [], [], []
    for level, c in enumerate(coefficients):
        c = scipy.sparse.coo_matrix(c)
        ts.append(np.asarray(c.row, dtype=np.int64))
        lv.append(np.full(c.nnz, level, dtype=np.int64))
        ks.append(np.asarray(c.col, dtype=np.int64))
        vs.append(np.asarray(c.data, dtype=np.float64))
    if not ts:
        return np.zeros(0, dtype=EVENT_DTYPE)
    t, l, k, v = np.concatenate(ts), np.concatenate(lv), np.concatenate(ks), np.concatenate(vs)
    order = np.argsort(t, kind='stable')
    events = np.empty(len(t), dtype=EVENT_DTYPE)
    events['f0'], events['f1'], events['f2'], events['f3'] = t[order], l[order], k[order], v[order]
    return events


def convertEventsToSparseMatrices(events, counts, sequenceLength):
    """hsc/dataset.py:813-824.  One csr_matrix [sequenceLength, counts[level]] per level (coefficient dtype = the events')."""
    events = np.asarray(events)
    t, lv, f, v = events['f0'].astype(int), events['f1'].astype(int), events['f2'].astype(int), events['f3']
    out = []
    for level, count in enumerate(counts):
        m = lv == level
        out.append(scipy.sparse.coo_matrix((v[m], (t[m], f[m])), shape=(sequenceLength, int(count))).tocsr())
    return out


def encodeResultToEvents(result, signal=0, level=0):
    """EncodeResult (the engine's per-signal event lists, duplicates allowed) -> the reference's event records for one
    level, duplicates summed like the accumulated code (hsc/modeling.py:992), sorted by time."""
    return convertSparseMatricesToEvents([scipy.sparse.csr_matrix((0, 0))] * level + [result.to_csc(signal, None)])
