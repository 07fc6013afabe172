"""Python host side of the engine: owns a native `hsc_engine` handle and uses PyTorch only for
device memory, pinned host buffers and streams.  Every number is produced by the CUDA kernels
behind the C ABI (include/hsc_b200.h); there is no CPU path in this module.
"""
import ctypes
import math

import numpy as np

from . import _native as N


def _torch():
    import torch
    return torch


def _np_dtype(code):
    return np.float32 if code == N.HSC_F32 else np.float64


def engine_dtype(*arrays):
    """Arithmetic type of the path: the NumPy result type of (signal, dictionary), as np.dot gives the
    reference (hsc/modeling.py:186-187): float32 only if everything is float32."""
    rt = np.result_type(*[np.asarray(a).dtype for a in arrays])
    return np.float32 if rt == np.float32 else np.float64


class EncodeResult(object):
    """Events of S signals in selection order, plus the final states and (optionally) residuals."""

    def __init__(self, S, T, K):
        self.S, self.T, self.K = S, T, K
        self.pos = [np.zeros(0, np.int32) for _ in range(S)]
        self.idx = [np.zeros(0, np.int32) for _ in range(S)]
        self.coef = [None] * S
        self.states = None
        self.residual = None
        self.counts = None          # events per signal (host pipeline; also set when the event lists stay on the device)
        self.gathered = None        # what the on_device_events hook of the host pipeline returned (multi-GPU gather)
        self.batch_index = None

    def stats(self, s=0):
        st = self.states[s]
        return dict(energy_signal=st.energy_signal, energy_residual=st.energy_residual, n_events=st.n_events, nnz=st.nnz,
                    duplicates=st.duplicates, passes=st.passes, reranked=st.reranked,
                    stop=N.STOP_NAMES.get(st.status, str(st.status)))

    def total_events(self):
        if self.counts is not None:
            return int(np.sum(self.counts))
        return int(sum(len(p) for p in self.pos))

    def to_csc(self, s=0, min_coefficients=1e-16):
        """Accumulates the events of signal s into the reference's return type: csc_matrix [T,K]
        float64, duplicates summed (`+=`, hsc/modeling.py:992), |c| < minCoefficients dropped
        (:1171-1177), zeros eliminated (:1180-1181)."""
        import scipy.sparse
        m = scipy.sparse.coo_matrix((self.coef[s].astype(np.float64), (self.pos[s].astype(np.int64), self.idx[s].astype(np.int64))),
                                    shape=(self.T, self.K)).tocsc()
        m.sum_duplicates()
        if min_coefficients is not None:
            m.data[np.abs(m.data) < min_coefficients] = 0.0
        m.eliminate_zeros()
        return m


def _with_stream(fn):
    """A `stream=` argument becomes torch's CURRENT stream for the duration of the call, so that the tensor allocations, the
    non-blocking host-to-device copies and the torch.distributed collectives inside are ordered with the native launches
    (which go to the same stream through _stream_ptr) instead of racing them from the default stream."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kw):
        stream = kw.get('stream')
        if stream is None:
            return fn(self, *args, **kw)
        with _torch().cuda.stream(stream):
            return fn(self, *args, **kw)
    return wrapper


class Engine(object):
    """One native engine on one CUDA device."""

    def __init__(self, device=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError('hierarchical_sparse_coding_b200 needs a CUDA device (no CPU fallback)')
        self.lib = N.load_library()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device('cuda', int(device) if not isinstance(device, torch.device) else device.index or 0)
        h = ctypes.c_void_p()
        N.check(self.lib, None, self.lib.hsc_b200_create(self.device.index, ctypes.byref(h)))
        self.handle = h
        self.dtype = None
        self.K = self.L = self.F = None
        self._D_host = None
        self._w_host = None

    def close(self):
        self._pipe_cache = self._host_cache = self._ws_cache = self._last_workspace = None    # device + pinned staging buffers
        for h in getattr(self, '_views', []):
            self.lib.hsc_b200_destroy(h)
        self._views = []
        if getattr(self, 'handle', None):
            self.lib.hsc_b200_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(self.lib.hsc_b200_launch_count(self.handle))

    # ------------------------------------------------------------------ dictionary
    def set_dictionary(self, D, weights=None, dtype=None):
        D = np.asarray(D)
        assert D.ndim == 2 or D.ndim == 3
        if D.ndim == 2:
            D = D[:, :, None]
        dt = np.dtype(dtype) if dtype is not None else np.dtype(engine_dtype(D))
        assert dt in (np.dtype(np.float32), np.dtype(np.float64))
        Dh = np.ascontiguousarray(D, dtype=dt)
        wh = None
        if weights is not None:
            wh = np.ascontiguousarray(np.asarray(weights), dtype=dt)
            assert wh.shape == (Dh.shape[0],)
        # the same dictionary again (every encode of a coder re-sends it): keep the uploaded D, Gram tensor and K1 operand
        if (self._D_host is not None and self.dtype == dt and self._D_host.shape == Dh.shape and np.array_equal(self._D_host, Dh)
                and ((wh is None and self._w_host is None) or
                     (wh is not None and self._w_host is not None and np.array_equal(self._w_host, wh)))):
            return self
        code = N.HSC_F32 if dt == np.dtype(np.float32) else N.HSC_F64
        with _torch().cuda.device(self.device):
            N.check(self.lib, self.handle, self.lib.hsc_b200_set_dictionary(
                self.handle, Dh.ctypes.data_as(ctypes.c_void_p), code, Dh.shape[0], Dh.shape[1], Dh.shape[2],
                wh.ctypes.data_as(ctypes.c_void_p) if wh is not None else None))
        self.dtype = dt
        self.K, self.L, self.F = Dh.shape
        self._dict_version = getattr(self, '_dict_version', 0) + 1
        self._D_host, self._w_host = Dh.copy(), (None if wh is None else wh.copy())
        return self

    @property
    def torch_dtype(self):
        torch = _torch()
        return torch.float32 if self.dtype == np.dtype(np.float32) else torch.float64

    def _to_host(self, ptr, shape):
        out = np.empty(shape, dtype=self.dtype)
        N.check(self.lib, self.handle, self.lib.hsc_b200_copy_to_host(
            self.handle, ctypes.c_void_p(ptr), out.ctypes.data_as(ctypes.c_void_p), out.nbytes))
        return out

    def gram(self):
        """Copy of the device Gram tensor G[K][2L-1][K] (tests)."""
        return self._to_host(self.lib.hsc_b200_gram_dev(self.handle), (self.K, 2 * self.L - 1, self.K))

    # ------------------------------------------------------------------ helpers
    def _stream_ptr(self, stream=None):
        torch = _torch()
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        return ctypes.c_void_p(s.cuda_stream)

    def _as_device_batch(self, x, stream=None):
        """x: numpy [S,T,F] / torch tensor (host or device) -> contiguous device tensor of the engine dtype."""
        torch = _torch()
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=self.dtype))
        x = x.to(device=self.device, dtype=self.torch_dtype, non_blocking=True).contiguous()
        assert x.dim() == 3 and x.shape[2] == self.F, 'signals must be [S,T,F=%d]' % self.F
        return x

    def make_options(self, nbNonzeroCoefs=None, toleranceResidualScale=None, toleranceSnr=None, nbBlocks=1,
                     minCoefficients=1e-16, use_weights=False, coef_mode=1, max_passes_per_run=0, max_events_total=0,
                     method=0, rerank_tolerance=-1.0, energy_eps=None):
        o = N.MpOptions()
        o.nb_nonzero_coefs = -1 if nbNonzeroCoefs is None else int(nbNonzeroCoefs)
        o.tolerance_snr = float('nan') if toleranceSnr is None else float(toleranceSnr)
        o.tolerance_residual_scale = float('nan') if toleranceResidualScale is None else float(toleranceResidualScale)
        o.min_coefficients = -1.0 if minCoefficients is None else float(minCoefficients)
        o.nb_blocks = -1 if nbBlocks == 'auto' else int(nbBlocks)
        o.use_weights = 1 if use_weights else 0
        o.coef_mode = int(coef_mode)
        o.max_passes_per_run = int(max_passes_per_run)
        o.max_events_total = int(max_events_total)
        o.method = int(method)
        o.rerank_tolerance = float(rerank_tolerance)       # < 0: the engine's default near-tie window (include/hsc_b200.h)
        o.energy_eps = 0.0 if energy_eps is None else float(energy_eps)     # np.finfo(D.dtype).eps of the caller's dictionary (:1057)
        return o

    # ------------------------------------------------------------------ correlation (K1)
    @_with_stream
    def correlate(self, x, stream=None):
        """convolve1d(x, D, 'same') for a batch: returns the device map [S,T,K]."""
        torch = _torch()
        with torch.cuda.device(self.device):
            xd = self._as_device_batch(x)
            S, T, _ = xd.shape
            out = torch.empty((S, T, self.K), dtype=self.torch_dtype, device=self.device)
            N.check(self.lib, self.handle, self.lib.hsc_b200_correlate(
                self.handle, ctypes.c_void_p(xd.data_ptr()), S, T, ctypes.c_void_p(out.data_ptr()), self._stream_ptr(stream)))
        return out

    # ------------------------------------------------------------------ pursuit (K1 + K2)
    def workspace_bytes(self, S, T):
        return int(self.lib.hsc_b200_workspace_bytes(self.handle, S, T))

    def max_signals_per_chunk(self, T, budget_bytes=None):
        torch = _torch()
        if budget_bytes is None:
            free, _ = torch.cuda.mem_get_info(self.device)
            # blocks torch's caching allocator holds but has not handed out are reusable too
            free += torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
            budget_bytes = int(free * 0.8)
        per = self.workspace_bytes(1, T) + 2 * T * self.F * self.dtype.itemsize
        return max(1, int(budget_bytes // per))

    def default_capacity(self, options, T):
        if options.method == 1:       # LoCOMP emits one event per refitted group atom (<= 64 per selection)
            base = int(options.nb_nonzero_coefs * 4) if options.nb_nonzero_coefs >= 0 else int(max(1024, T // 8))
            return int(min(max(2048, base + 1024), 1 << 20))
        if options.max_events_total > 0:
            return int(options.max_events_total)
        if options.nb_nonzero_coefs >= 0:
            return int(options.nb_nonzero_coefs * 1.5) + 64
        return int(min(max(1024, T // 8), 1 << 20))

    @_with_stream
    def encode(self, x, options, capacity=None, return_residual=True, residual_inplace=False, stream=None,
               on_pass=None):
        """Matching pursuit of S independent signals.

        x: [S,T,F] numpy array, host tensor (ideally pinned) or device tensor.
        Returns EncodeResult with numpy event lists; `residual` is a device tensor [S,T,F] when
        return_residual (D2H is the caller's choice: .cpu()).
        on_pass(result_so_far, states) -> bool, if given, is called between launches
        (options.max_passes_per_run should be 1 then); returning True stops the encode."""
        torch = _torch()
        with torch.cuda.device(self.device):
            xd = self._as_device_batch(x)
            S, T, _ = xd.shape
            res = EncodeResult(S, T, self.K)
            if S == 0:
                return res
            cap = int(capacity) if capacity is not None else self.default_capacity(options, T)
            wsb = self.workspace_bytes(S, T)
            ws = torch.empty((wsb,), dtype=torch.uint8, device=self.device)
            resid = xd if residual_inplace else torch.empty_like(xd)
            evp = torch.empty((S, cap), dtype=torch.int32, device=self.device)
            evi = torch.empty((S, cap), dtype=torch.int32, device=self.device)
            evc = torch.empty((S, cap), dtype=self.torch_dtype, device=self.device)
            sp = self._stream_ptr(stream)
            N.check(self.lib, self.handle, self.lib.hsc_b200_mp_begin(
                self.handle, ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(resid.data_ptr()), S, T,
                ctypes.c_void_p(ws.data_ptr()), wsb, ctypes.byref(options), sp))
            states = (N.SignalState * S)()
            chunks_p = [[] for _ in range(S)]
            chunks_i = [[] for _ in range(S)]
            chunks_c = [[] for _ in range(S)]
            while True:
                N.check(self.lib, self.handle, self.lib.hsc_b200_mp_run(
                    self.handle, ctypes.c_void_p(evp.data_ptr()), ctypes.c_void_p(evi.data_ptr()),
                    ctypes.c_void_p(evc.data_ptr()), cap, states, sp))
                nb = np.array([states[s].n_buffered for s in range(S)], dtype=np.int64)
                m = int(nb.max())
                if m > 0:
                    hp = evp[:, :m].cpu().numpy()
                    hi = evi[:, :m].cpu().numpy()
                    hc = evc[:, :m].cpu().numpy()
                    for s in range(S):
                        if nb[s] > 0:
                            chunks_p[s].append(hp[s, :nb[s]].copy())
                            chunks_i[s].append(hi[s, :nb[s]].copy())
                            chunks_c[s].append(hc[s, :nb[s]].copy())
                paused = [states[s].status in (N.HSC_PAUSE_CAPACITY, N.HSC_PAUSE_PASSES, N.HSC_RUNNING) for s in range(S)]
                if not any(paused):
                    break
                if on_pass is not None:
                    self._fill(res, chunks_p, chunks_i, chunks_c, states)
                    res.residual = resid
                    if on_pass(res, states):
                        break
            self._fill(res, chunks_p, chunks_i, chunks_c, states)
            res.residual = resid if return_residual else None
            self._last_workspace = ws   # keeps the map alive for map_snapshot()
            self._last_shape = (S, T)
        return res

    @_with_stream
    def encode_device(self, xd, options, capacity, resid=None, stream=None, sync_states=True):
        """Lean path for resident data: xd is a device tensor [S,T,F] of the engine dtype; runs K1 + one
        K2 launch and returns (ev_pos, ev_idx, ev_coef, states, residual) with the events still on the
        device ([S,capacity] each).  `states` is None when sync_states is False (fully asynchronous)."""
        torch = _torch()
        S, T, _ = xd.shape
        with torch.cuda.device(self.device):
            wsb = self.workspace_bytes(S, T)
            ws = getattr(self, '_ws_cache', None)
            if ws is None or ws.numel() < wsb:
                self._ws_cache = None
                ws = torch.empty((wsb,), dtype=torch.uint8, device=self.device)
                self._ws_cache = ws
            if resid is None:
                resid = torch.empty_like(xd)
            evp = torch.empty((S, capacity), dtype=torch.int32, device=self.device)
            evi = torch.empty((S, capacity), dtype=torch.int32, device=self.device)
            evc = torch.empty((S, capacity), dtype=self.torch_dtype, device=self.device)
            sp = self._stream_ptr(stream)
            N.check(self.lib, self.handle, self.lib.hsc_b200_mp_begin(
                self.handle, ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(resid.data_ptr()), S, T,
                ctypes.c_void_p(ws.data_ptr()), wsb, ctypes.byref(options), sp))
            states = (N.SignalState * S)() if sync_states else None
            N.check(self.lib, self.handle, self.lib.hsc_b200_mp_run(
                self.handle, ctypes.c_void_p(evp.data_ptr()), ctypes.c_void_p(evi.data_ptr()),
                ctypes.c_void_p(evc.data_ptr()), capacity, states, sp))
            self._last_workspace = ws
            self._last_shape = (S, T)
        return evp, evi, evc, states, resid

    @_with_stream
    def begin_only(self, xd, options, resid, stream=None):
        """K1 + state reset only (bench: times the correlation separately from the pursuit)."""
        torch = _torch()
        S, T, _ = xd.shape
        with torch.cuda.device(self.device):
            wsb = self.workspace_bytes(S, T)
            ws = getattr(self, '_ws_cache', None)
            if ws is None or ws.numel() < wsb:
                self._ws_cache = None
                ws = torch.empty((wsb,), dtype=torch.uint8, device=self.device)
                self._ws_cache = ws
            N.check(self.lib, self.handle, self.lib.hsc_b200_mp_begin(
                self.handle, ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(resid.data_ptr()), S, T,
                ctypes.c_void_p(ws.data_ptr()), wsb, ctypes.byref(options), self._stream_ptr(stream)))
            self._last_workspace = ws
            self._last_shape = (S, T)

    @_with_stream
    def run_only(self, evp, evi, evc, capacity, sync_states=True, stream=None):
        S = self._last_shape[0]
        states = (N.SignalState * S)() if sync_states else None
        with _torch().cuda.device(self.device):
            N.check(self.lib, self.handle, self.lib.hsc_b200_mp_run(
                self.handle, ctypes.c_void_p(evp.data_ptr()), ctypes.c_void_p(evi.data_ptr()),
                ctypes.c_void_p(evc.data_ptr()), capacity, states, self._stream_ptr(stream)))
        return states

    @_with_stream
    def events_to_dense(self, evp, evi, evc, min_coefficients=1e-16, stream=None):
        """Level hand-off of the hierarchical encoder on the device (hsc/modeling.py:1489): the accumulated code of the
        encode in flight as a dense float64 device tensor [S,T,K] - the next level's K-channel input."""
        torch = _torch()
        S, T = self._last_shape
        with torch.cuda.device(self.device):
            out = torch.empty((S, T, self.K), dtype=torch.float64, device=self.device)
            N.check(self.lib, self.handle, self.lib.hsc_b200_mp_events_to_dense(
                self.handle, ctypes.c_void_p(evp.data_ptr()), ctypes.c_void_p(evi.data_ptr()), ctypes.c_void_p(evc.data_ptr()),
                int(evp.shape[1]), -1.0 if min_coefficients is None else float(min_coefficients), ctypes.c_void_p(out.data_ptr()),
                self._stream_ptr(stream)))
        return out

    # ------------------------------------------------------------------ resident pipeline (several encodes in flight)
    def make_slots(self, n, S, T, capacity):
        """n independent encode slots on this engine's dictionary (native views: own workspace, event buffers, residual,
        pinned states), so that the correlation of one batch can run - on its own stream - while the pursuit of the
        previous one is still going, and the pursuit CTAs of the next batch move into the SMs the previous batch's tail
        frees.  See EncodeSlot."""
        views = self._views_for(n)
        return [EncodeSlot(self, views[i], S, T, capacity) for i in range(n)]

    @_with_stream
    def compact_events(self, evp, evi, evc, out=None, stream=None):
        """Compacts the [S,capacity] event buffers of the encode in flight (after run_only / encode_device) on the device:
        returns dict(offsets=int64[S+1], pos, idx, coef flat device tensors of S*capacity entries); the atoms of signal s are
        [offsets[s], offsets[s+1]).  `out`: a dict from an earlier call to reuse its buffers.  Asynchronous."""
        torch = _torch()
        S, cap = evp.shape
        with torch.cuda.device(self.device):
            if out is None or out['pos'].numel() < S * cap or out['offsets'].numel() != S + 1:
                out = dict(offsets=torch.empty((S + 1,), dtype=torch.int64, device=self.device),
                           pos=torch.empty((S * cap,), dtype=torch.int32, device=self.device),
                           idx=torch.empty((S * cap,), dtype=torch.int32, device=self.device),
                           coef=torch.empty((S * cap,), dtype=self.torch_dtype, device=self.device))
            N.check(self.lib, self.handle, self.lib.hsc_b200_mp_compact_events(
                self.handle, ctypes.c_void_p(evp.data_ptr()), ctypes.c_void_p(evi.data_ptr()), ctypes.c_void_p(evc.data_ptr()), cap,
                ctypes.c_void_p(out['offsets'].data_ptr()), ctypes.c_void_p(out['pos'].data_ptr()), ctypes.c_void_p(out['idx'].data_ptr()),
                ctypes.c_void_p(out['coef'].data_ptr()), S * cap, self._stream_ptr(stream)))
        return out

    # ------------------------------------------------------------------ host pipeline (public batched path)
    def _views_for(self, n):
        """n extra native handles sharing this engine's device dictionary, one per in-flight chunk."""
        ver = getattr(self, '_dict_version', 0)
        if getattr(self, '_views_version', None) != ver or len(getattr(self, '_views', [])) < n:
            for h in getattr(self, '_views', []):
                self.lib.hsc_b200_destroy(h)
            self._views = []
            for _ in range(n):
                h = ctypes.c_void_p()
                N.check(self.lib, self.handle, self.lib.hsc_b200_create_view(self.handle, ctypes.byref(h)))
                self._views.append(h)
            self._views_version = ver
            self._chunk_cache = {}
        return self._views[:n]

    def encode_host(self, x_host, options, capacity=None, n_chunks=4, residual_out=None, want_residual=True):
        """Matching pursuit of S independent signals that live in HOST memory (ideally a pinned torch tensor
        [S,T,F] of the engine dtype).  The host-to-device copy is cut into `n_chunks` chunks on a copy stream;
        the correlation (K1) of a chunk starts as soon as its copy has landed (hsc_b200_mp_begin_part), so the
        PCIe transfer hides under K1.  The select/update loop (K2) then runs once over the whole batch - it
        needs every signal resident to fill the GPU - and codes + residuals are copied back.
        Residuals land in `residual_out` (host tensor like x; allocated, pinned if x is, when None and wanted).
        Returns EncodeResult (events as numpy arrays, residual = the host tensor)."""
        torch = _torch()
        if isinstance(x_host, np.ndarray):
            x_host = torch.from_numpy(np.ascontiguousarray(x_host, dtype=self.dtype))
        assert x_host.dim() == 3 and x_host.shape[2] == self.F and x_host.dtype == self.torch_dtype
        S, T, _ = x_host.shape
        res = EncodeResult(S, T, self.K)
        if S == 0:
            return res
        cap = int(capacity) if capacity is not None else self.default_capacity(options, T)
        n_chunks = max(1, min(int(n_chunks), S))
        if want_residual and residual_out is None:
            residual_out = torch.empty_like(x_host).pin_memory() if x_host.is_pinned() else torch.empty_like(x_host)
        with torch.cuda.device(self.device):
            key = (S, T, cap, self.F, str(self.dtype))
            cache = getattr(self, '_host_cache', None)
            if cache is None or cache['key'] != key:
                self._host_cache = None
                wsb = self.workspace_bytes(S, T)
                cache = dict(key=key, wsb=wsb, copy_stream=torch.cuda.Stream(device=self.device),
                             ws=torch.empty((wsb,), dtype=torch.uint8, device=self.device),
                             xd=torch.empty((S, T, self.F), dtype=self.torch_dtype, device=self.device),
                             evp=torch.empty((S, cap), dtype=torch.int32, device=self.device),
                             evi=torch.empty((S, cap), dtype=torch.int32, device=self.device),
                             evc=torch.empty((S, cap), dtype=self.torch_dtype, device=self.device),
                             hp=torch.empty((S, cap), dtype=torch.int32).pin_memory(),
                             hi=torch.empty((S, cap), dtype=torch.int32).pin_memory(),
                             hc=torch.empty((S, cap), dtype=self.torch_dtype).pin_memory(),
                             states=(N.SignalState * S)())
                self._host_cache = cache
            cur = torch.cuda.current_stream(self.device)
            cs = cache['copy_stream']
            cs.wait_stream(cur)                       # the staging buffer may still be read by earlier work
            sp = ctypes.c_void_p(cur.cuda_stream)
            xp = ctypes.c_void_p(cache['xd'].data_ptr())
            wsp = ctypes.c_void_p(cache['ws'].data_ptr())
            for c in range(n_chunks):
                lo, hi = S * c // n_chunks, S * (c + 1) // n_chunks
                with torch.cuda.stream(cs):
                    cache['xd'][lo:hi].copy_(x_host[lo:hi], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                cur.wait_event(ev)
                N.check(self.lib, self.handle, self.lib.hsc_b200_mp_begin_part(
                    self.handle, xp, xp, S, T, wsp, cache['wsb'], ctypes.byref(options), lo, hi - lo, sp))
            stt = cache['states']
            pp, ip, cp_ = (ctypes.c_void_p(cache['evp'].data_ptr()), ctypes.c_void_p(cache['evi'].data_ptr()),
                           ctypes.c_void_p(cache['evc'].data_ptr()))
            chunks_p = [[] for _ in range(S)]
            chunks_i = [[] for _ in range(S)]
            chunks_c = [[] for _ in range(S)]
            while True:
                N.check(self.lib, self.handle, self.lib.hsc_b200_mp_run(self.handle, pp, ip, cp_, cap, None, sp))
                N.check(self.lib, self.handle, self.lib.hsc_b200_mp_states_async(self.handle, stt, sp))
                cache['hp'].copy_(cache['evp'], non_blocking=True)
                cache['hi'].copy_(cache['evi'], non_blocking=True)
                cache['hc'].copy_(cache['evc'], non_blocking=True)
                cur.synchronize()
                hp, hi_, hc = cache['hp'].numpy(), cache['hi'].numpy(), cache['hc'].numpy()
                for s in range(S):
                    nb = stt[s].n_buffered
                    if nb > 0:
                        chunks_p[s].append(hp[s, :nb].copy())
                        chunks_i[s].append(hi_[s, :nb].copy())
                        chunks_c[s].append(hc[s, :nb].copy())
                if not any(stt[s].status in (N.HSC_PAUSE_CAPACITY, N.HSC_PAUSE_PASSES, N.HSC_RUNNING) for s in range(S)):
                    break
            if want_residual:
                residual_out.copy_(cache['xd'], non_blocking=True)
                cur.synchronize()
            self._fill(res, chunks_p, chunks_i, chunks_c, stt)
            res.residual = residual_out if want_residual else None
            self._last_workspace = cache['ws']
            self._last_shape = (S, T)
        return res

    # ------------------------------------------------------------------ K-SVD dictionary update
    def accumulate_code(self, sig, pos, idx, coef, S, T, K, min_coefficients=1e-16):
        """Event lists (selection order, duplicates allowed) -> the accumulated code the reference returns
        (`+=` per (t,k), hsc/modeling.py:992; |c| < minCoefficients dropped, :1171-1177; zeros eliminated, :1181),
        as device tensors sorted by (filter, signal, position) plus the per-filter column pointer (host, [K+1]).
        Sorting / compaction is torch tensor plumbing; the arithmetic on the code happens in the engine."""
        torch = _torch()
        with torch.cuda.device(self.device):
            to = lambda a, dt: (a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))).to(self.device, dt)
            sig, pos, idx, coef = to(sig, torch.int64), to(pos, torch.int64), to(idx, torch.int64), to(coef, torch.float64)
            key = (idx * int(S) + sig) * int(T) + pos
            key, order = torch.sort(key, stable=True)
            uniq, inv = torch.unique_consecutive(key, return_inverse=True)
            c = torch.zeros(uniq.numel(), dtype=torch.float64, device=self.device).index_add_(0, inv, coef[order])
            keep = c != 0.0
            if min_coefficients is not None:
                keep &= c.abs() >= float(min_coefficients)
            uniq, c = uniq[keep], c[keep].contiguous()
            p = (uniq % int(T)).to(torch.int32).contiguous()
            rest = uniq // int(T)
            sg = (rest % int(S)).to(torch.int32).contiguous()
            ix = (rest // int(S)).to(torch.int32).contiguous()
            counts = torch.bincount(ix.long(), minlength=int(K)).cpu().numpy().astype(np.int64)
            col_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        return sg, p, ix, c, col_ptr

    @_with_stream
    def ksvd_update(self, D, sig, pos, idx, coef, col_ptr, S, T, stream=None, group=None, use_pca=False):
        """One dictionary-update stage (hsc/modeling.py:593-636) on the device, float64.  D: numpy [K,L,F];
        (sig, pos, idx, coef, col_ptr) as accumulate_code returns them.  Returns (D_new numpy float64,
        coef_new device tensor, alpha).

        `group`: a torch.distributed process group (or True for the default group) when every rank holds the code of
        its own signals and the same D: per filter the q x q window Gram matrices are summed over the ranks
        (all_reduce, the one exchange the dictionary update needs) and every rank derives the same filter.
        `use_pca`: the usePCA=True variant (:618-625): first principal component of the mean-centred windows."""
        torch = _torch()
        if use_pca and group is not None:
            raise NotImplementedError('usePCA=True under a process group: the window means would have to be shared too')
        N.check(self.lib, self.handle, self.lib.hsc_b200_ksvd_set_pca(self.handle, 1 if use_pca else 0))
        D = np.ascontiguousarray(D, dtype=np.float64)
        K, L, F = D.shape
        with torch.cuda.device(self.device):
            Dd = torch.from_numpy(D).to(self.device)
            coef = coef.clone()
            cp = np.ascontiguousarray(col_ptr, dtype=np.int64)
            assert cp.shape == (K + 1,) and int(cp[-1]) == int(coef.numel())
            alpha = ctypes.c_double(0.0)
            cpp = cp.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
            args = (ctypes.c_void_p(Dd.data_ptr()), K, L, F, cpp, ctypes.c_void_p(sig.data_ptr()), ctypes.c_void_p(pos.data_ptr()),
                    ctypes.c_void_p(idx.data_ptr()), ctypes.c_void_p(coef.data_ptr()), int(S), int(T))
            if group is None:
                N.check(self.lib, self.handle, self.lib.hsc_b200_ksvd_update(self.handle, *args, ctypes.byref(alpha), self._stream_ptr(stream)))
                return Dd.cpu().numpy(), coef, float(alpha.value)
            import torch.distributed as dist
            pg = None if group is True else group
            counts = torch.from_numpy(np.diff(cp)).to(self.device)
            dist.all_reduce(counts, group=pg)                      # which filters have an atom on ANY rank (:598-599)
            counts = counts.cpu().numpy()
            C = torch.empty((L * F, L * F), dtype=torch.float64, device=self.device)
            sweep = ctypes.c_void_p()
            N.check(self.lib, self.handle, self.lib.hsc_b200_ksvd_begin(
                self.handle, *args, ctypes.c_void_p(C.data_ptr()), self._stream_ptr(stream), ctypes.byref(sweep)))
            try:
                for k in range(K):
                    if counts[k] == 0:
                        continue
                    N.check(self.lib, self.handle, self.lib.hsc_b200_ksvd_filter_gram(sweep, k, None))
                    dist.all_reduce(C, group=pg)                   # sum of the ranks' window Gram matrices
                    N.check(self.lib, self.handle, self.lib.hsc_b200_ksvd_filter_finish(sweep, k, 0))
            finally:
                rc = self.lib.hsc_b200_ksvd_end(sweep, ctypes.byref(alpha))
            N.check(self.lib, self.handle, rc)
            return Dd.cpu().numpy(), coef, float(alpha.value)

    def encode_chunked(self, x, options, capacity=None, budget_bytes=None):
        """encode() over as many signals at a time as the correlation maps fit in free HBM.  x: numpy [S,T,F].
        Returns one EncodeResult for all S signals (no residual)."""
        S, T, _ = x.shape
        per = self.max_signals_per_chunk(T, budget_bytes)
        out = EncodeResult(S, T, self.K)
        out.states = []
        for lo in range(0, S, per):
            r = self.encode(x[lo:lo + per], options, capacity=capacity, return_residual=False)
            n = r.S
            out.pos[lo:lo + n], out.idx[lo:lo + n], out.coef[lo:lo + n] = r.pos, r.idx, r.coef
            out.states.extend(r.states)
            self._last_workspace = None
        return out

    # ------------------------------------------------------------------ convolutional k-means assignment step
    @_with_stream
    def kmeans_assign(self, windows, stream=None):
        """Assignment step of the convolutional k-means learner (hsc/modeling.py:455-480) with the centroids set as
        the dictionary.  windows: device tensor [B,Tw,F] of the engine dtype.  Returns (pos[B] int32 = first sample
        of each window's best patch, idx[B] int32 = its centroid, sums[K,L,F] float64 = sum of the L2-normalised
        assigned patches, counts[K] int32), all device tensors."""
        torch = _torch()
        B, Tw, _ = windows.shape
        with torch.cuda.device(self.device):
            scratch = torch.empty((B, Tw, self.K), dtype=self.torch_dtype, device=self.device)
            pos = torch.empty((B,), dtype=torch.int32, device=self.device)
            idx = torch.empty((B,), dtype=torch.int32, device=self.device)
            sums = torch.empty((self.K, self.L, self.F), dtype=torch.float64, device=self.device)
            counts = torch.empty((self.K,), dtype=torch.int32, device=self.device)
            N.check(self.lib, self.handle, self.lib.hsc_b200_kmeans_assign(
                self.handle, ctypes.c_void_p(windows.data_ptr()), int(B), int(Tw), ctypes.c_void_p(scratch.data_ptr()),
                ctypes.c_void_p(pos.data_ptr()), ctypes.c_void_p(idx.data_ptr()), ctypes.c_void_p(sums.data_ptr()),
                ctypes.c_void_p(counts.data_ptr()), self._stream_ptr(stream)))
        return pos, idx, sums, counts

    def encode_host_pipelined(self, batches, options, capacity=None, n_chunks=8, want_residual=False, residual_outs=None,
                              host_events=True, on_device_events=None, two_workspaces=True):
        """Generator over EncodeResult, one per batch of `batches` (an iterable of host tensors [S,T,F] of the engine
        dtype, ideally pinned, all the same shape).  Per batch: chunked host-to-device copy hidden under the correlation
        (hsc_b200_mp_begin_part), the select/update loop, compaction of the event buffers on the device
        (hsc_b200_mp_compact_events) and the device-to-host read of the codes - exactly 12-16 bytes per atom.  Two sets of
        staging buffers let the copies out of batch i (their own stream) overlap the copy in and the correlation of batch
        i+1 (PCIe is full duplex), and the host-side unpacking of batch i overlaps the GPU work of batch i+1.

        want_residual: also copy the residual [S,T,F] back (as large as the input; off by default - the codes are the
          result, the residual is x - decode(codes)).  `residual_outs`: one host tensor per batch to receive it.
        host_events=False: skip the device-to-host read of the codes (ranks that only feed a gather); the result then
          carries the per-signal counts / states only.
        on_device_events(batch_index, dev): called when a batch is finished, under the code-copy stream, with dev =
          dict(offsets=int64[S+1], pos=int32[..], idx=int32[..], coef=[..], total=n) - the compacted codes still on the
          device (valid until the hook's stream work is done): the hook for the NCCL gather of the sparse codes
          (distributed.gather_device_events); its return value becomes result.gathered.
        two_workspaces: keep one workspace per staging slot when device memory allows (2 x the correlation maps), so that
          the correlation of batch i+1 overlaps the tail of batch i's pursuit.
        A batch whose event buffers overflow (`capacity`) raises: use encode_host for open-ended stop rules."""
        torch = _torch()
        it = iter(batches)
        # staging buffers (device + pinned host) and streams are kept on the engine between calls: pinned allocations
        # cost tens of milliseconds
        cache = getattr(self, '_pipe_cache', None)
        if cache is None:
            cache = self._pipe_cache = dict(slots=[None, None], ctx=dict(copy_in=None, copy_out=None))
        slots, ctx = cache['slots'], cache['ctx']
        for sl_ in slots:                     # an abandoned earlier call may have left copies in flight
            if sl_ is not None and sl_['d2h_done'] is not None:
                sl_['d2h_done'].synchronize()
                sl_['d2h_done'] = None
        ctx['states_copied'] = None
        pending = None          # (slot index, EncodeResult shell, residual_out)
        sz_state = ctypes.sizeof(N.SignalState)

        def make_slot(S, T, cap):
            n_max = S * cap
            d = dict(xd=torch.empty((S, T, self.F), dtype=self.torch_dtype, device=self.device),
                     evp=torch.empty((S, cap), dtype=torch.int32, device=self.device),
                     evi=torch.empty((S, cap), dtype=torch.int32, device=self.device),
                     evc=torch.empty((S, cap), dtype=self.torch_dtype, device=self.device),
                     offsets=torch.empty((S + 1,), dtype=torch.int64, device=self.device),
                     cpos=torch.empty((n_max,), dtype=torch.int32, device=self.device),
                     cidx=torch.empty((n_max,), dtype=torch.int32, device=self.device),
                     ccoef=torch.empty((n_max,), dtype=self.torch_dtype, device=self.device),
                     hoff=torch.empty((S + 1,), dtype=torch.int64).pin_memory(),
                     hp=torch.empty((n_max,), dtype=torch.int32).pin_memory(),
                     hi=torch.empty((n_max,), dtype=torch.int32).pin_memory(),
                     hc=torch.empty((n_max,), dtype=self.torch_dtype).pin_memory(),
                     hstate=torch.empty((S * sz_state,), dtype=torch.uint8).pin_memory(),
                     d2h_done=None)
            d['states'] = (N.SignalState * S).from_address(d['hstate'].data_ptr())
            return d

        def finish(p):
            k, res, residual_out = p
            sl = slots[k]
            cout = ctx['copy_codes']                  # its own stream: copy_out already holds the next batch's (waiting) work
            sl['small_done'].synchronize()            # states + offsets are on the host
            S = res.S
            # one private copy of the S states, then vectorised unpacking (a Python loop over 512 signals with three
            # array copies each costs several milliseconds per batch)
            snap = (N.SignalState * S).from_buffer_copy(bytes(sl['hstate'].numpy()[:S * sz_state]))
            raw = np.frombuffer(snap, dtype=np.uint8).reshape(S, sz_state)
            status = raw[:, N.SignalState.status.offset:N.SignalState.status.offset + 4].copy().view(np.int32)[:, 0]
            if np.any((status == N.HSC_PAUSE_CAPACITY) | (status == N.HSC_PAUSE_PASSES) | (status == N.HSC_RUNNING)):
                sl['d2h_done'].synchronize()
                raise N.HscError(N.HSC_E_NOMEM, 'encode_host_pipelined: event capacity exhausted before the stop rule fired')
            sl['d2h_done'].synchronize()              # residual (if wanted) and the gather hook's reads of this slot
            off = sl['hoff'].numpy().copy()
            total = int(off[-1])
            res.counts = np.diff(off)
            if on_device_events is not None:
                # the sparse codes are still on the device, compacted: the hook gathers them over the ranks (NCCL)
                with torch.cuda.stream(cout):
                    res.gathered = on_device_events(res.batch_index, dict(offsets=sl['offsets'], pos=sl['cpos'], idx=sl['cidx'],
                                                                           coef=sl['ccoef'], total=total))
            if host_events:
                # the codes: exactly `total` atoms cross the bus
                with torch.cuda.stream(cout):
                    sl['hp'][:total].copy_(sl['cpos'][:total], non_blocking=True)
                    sl['hi'][:total].copy_(sl['cidx'][:total], non_blocking=True)
                    sl['hc'][:total].copy_(sl['ccoef'][:total], non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(cout)
                sl['d2h_done'] = done
                done.synchronize()
                cuts = off[1:-1]
                res.pos = np.split(sl['hp'].numpy()[:total].copy(), cuts)
                res.idx = np.split(sl['hi'].numpy()[:total].copy(), cuts)
                res.coef = np.split(sl['hc'].numpy()[:total].copy(), cuts)
            res.states = [snap[i] for i in range(S)]
            res.residual = residual_out
            return res

        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            for bi, x_host in enumerate(it):
                if isinstance(x_host, np.ndarray):
                    x_host = torch.from_numpy(np.ascontiguousarray(x_host, dtype=self.dtype))
                assert x_host.dim() == 3 and x_host.shape[2] == self.F and x_host.dtype == self.torch_dtype
                S, T, _ = x_host.shape
                cap = int(capacity) if capacity is not None else self.default_capacity(options, T)
                k = bi & 1
                if slots[k] is None or slots[k]['xd'].shape != (S, T, self.F) or slots[k]['evp'].shape[1] != cap:
                    slots[k] = make_slot(S, T, cap)
                wsb = self.workspace_bytes(S, T)
                if ctx.get('copy_in') is None:
                    ctx['copy_in'] = torch.cuda.Stream(device=self.device)
                    ctx['copy_out'] = torch.cuda.Stream(device=self.device)
                    ctx['copy_codes'] = torch.cuda.Stream(device=self.device)
                    ctx['k1'] = torch.cuda.Stream(device=self.device, priority=-1)
                    ctx['k2'] = [torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)]
                    ctx['wss'] = [None, None]
                # Two workspaces (one per staging slot) when they fit: the correlation of batch i+1 - on its own stream,
                # under the chunks of its copy in - then starts while the pursuit of batch i is still in its tail, and the
                # pursuit of batch i+1 moves into the SMs that tail frees.  One shared workspace otherwise: the correlation
                # of batch i+1 waits for the pursuit of batch i.
                if ctx['wss'][k] is None or ctx['wss'][k].numel() < wsb:
                    ctx['wss'][k] = None
                    other = ctx['wss'][k ^ 1]
                    free, _ = torch.cuda.mem_get_info(self.device)
                    free += torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
                    if two_workspaces and (other is None or other.numel() < wsb or free > wsb + (4 << 30)):
                        ctx['wss'][k] = torch.empty((wsb,), dtype=torch.uint8, device=self.device)
                    else:
                        if other is None or other.numel() < wsb:
                            ctx['wss'][k ^ 1] = other = torch.empty((wsb,), dtype=torch.uint8, device=self.device)
                        ctx['wss'][k] = other                       # shared
                shared_ws = ctx['wss'][0] is ctx['wss'][1]
                handles = [self.handle, self.handle] if shared_ws else self._views_for(2)
                sl, ws, cin, cout = slots[k], ctx['wss'][k], ctx['copy_in'], ctx['copy_out']
                s_k1, s_k2, hk = ctx['k1'], ctx['k2'][k], handles[k]
                if bi == 0:
                    for st_ in (cin, s_k1, ctx['k2'][0], ctx['k2'][1]):
                        st_.wait_stream(cur)                        # work queued by the caller before this call
                residual_out = None
                if want_residual:
                    residual_out = residual_outs[bi] if residual_outs is not None else (
                        torch.empty_like(x_host).pin_memory() if x_host.is_pinned() else torch.empty_like(x_host))
                # this slot's staging buffers (and its workspace) are free once the copies of batch bi-2 have left them;
                # a shared workspace also waits for the pursuit and the compaction of batch bi-1
                if sl['d2h_done'] is not None:
                    cin.wait_event(sl['d2h_done'])
                    s_k1.wait_event(sl['d2h_done'])
                if shared_ws and ctx.get('states_copied') is not None:
                    s_k1.wait_event(ctx['states_copied'])
                xp, wsp = ctypes.c_void_p(sl['xd'].data_ptr()), ctypes.c_void_p(ws.data_ptr())
                k1p = ctypes.c_void_p(s_k1.cuda_stream)
                nch = max(1, min(int(n_chunks), S))
                for c in range(nch):
                    lo, hi = S * c // nch, S * (c + 1) // nch
                    with torch.cuda.stream(cin):
                        sl['xd'][lo:hi].copy_(x_host[lo:hi], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(cin)
                    s_k1.wait_event(ev)
                    N.check(self.lib, hk, self.lib.hsc_b200_mp_begin_part(
                        hk, xp, xp, S, T, wsp, wsb, ctypes.byref(options), lo, hi - lo, k1p))
                k1_done = torch.cuda.Event()
                k1_done.record(s_k1)
                s_k2.wait_event(k1_done)
                N.check(self.lib, hk, self.lib.hsc_b200_mp_run(
                    hk, ctypes.c_void_p(sl['evp'].data_ptr()), ctypes.c_void_p(sl['evi'].data_ptr()),
                    ctypes.c_void_p(sl['evc'].data_ptr()), cap, None, ctypes.c_void_p(s_k2.cuda_stream)))
                k2_done = torch.cuda.Event()
                k2_done.record(s_k2)
                if shared_ws:
                    s_k1.wait_event(k2_done)                        # the next correlation overwrites this map
                with torch.cuda.stream(cout):
                    cout.wait_event(k2_done)
                    csp = ctypes.c_void_p(cout.cuda_stream)
                    N.check(self.lib, hk, self.lib.hsc_b200_mp_states_async(hk, sl['states'], csp))
                    N.check(self.lib, hk, self.lib.hsc_b200_mp_compact_events(
                        hk, ctypes.c_void_p(sl['evp'].data_ptr()), ctypes.c_void_p(sl['evi'].data_ptr()),
                        ctypes.c_void_p(sl['evc'].data_ptr()), cap, ctypes.c_void_p(sl['offsets'].data_ptr()),
                        ctypes.c_void_p(sl['cpos'].data_ptr()), ctypes.c_void_p(sl['cidx'].data_ptr()),
                        ctypes.c_void_p(sl['ccoef'].data_ptr()), S * cap, csp))
                    ctx['states_copied'] = torch.cuda.Event()       # a shared workspace's states may be reset by the next batch
                    ctx['states_copied'].record(cout)
                    sl['hoff'].copy_(sl['offsets'], non_blocking=True)
                    small = torch.cuda.Event()
                    small.record(cout)
                    sl['small_done'] = small
                    if want_residual:
                        residual_out.copy_(sl['xd'], non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(cout)
                sl['d2h_done'] = done
                self._last_workspace, self._last_shape = ws, (S, T)
                shell = EncodeResult(S, T, self.K)
                shell.batch_index = bi
                if pending is not None:
                    yield finish(pending)
                pending = (k, shell, residual_out)
            if pending is not None:
                yield finish(pending)
            for st_ in (ctx['copy_out'], ctx['copy_codes']):
                cur.wait_stream(st_)

    def _fill(self, res, cp, ci, cc, states):
        for s in range(res.S):
            res.pos[s] = np.concatenate(cp[s]) if cp[s] else np.zeros(0, np.int32)
            res.idx[s] = np.concatenate(ci[s]) if ci[s] else np.zeros(0, np.int32)
            res.coef[s] = np.concatenate(cc[s]) if cc[s] else np.zeros(0, self.dtype)
        res.states = [N.SignalState.from_buffer_copy(bytes(states[s])) for s in range(res.S)]

    def map_snapshot(self):
        """Copy of the correlation map of the last encode [S,T,K] (tests / diagnostics)."""
        S, T = self._last_shape
        return self._to_host(self.lib.hsc_b200_mp_map_dev(self.handle), (S, T, self.K))

    # ------------------------------------------------------------------ decoder
    @_with_stream
    def decode(self, pos, idx, coef, T, out=None, stream=None):
        """reconstructSignal for one signal: returns the device tensor [T,F] (+= into `out` if given)."""
        torch = _torch()
        with torch.cuda.device(self.device):
            pos = np.asarray(pos, dtype=np.int32)
            order = np.argsort(pos, kind='stable')
            p = torch.from_numpy(np.ascontiguousarray(pos[order])).to(self.device)
            i = torch.from_numpy(np.ascontiguousarray(np.asarray(idx, dtype=np.int32)[order])).to(self.device)
            c = torch.from_numpy(np.ascontiguousarray(np.asarray(coef, dtype=self.dtype)[order])).to(self.device)
            if out is None:
                out = torch.zeros((T, self.F), dtype=self.torch_dtype, device=self.device)
            N.check(self.lib, self.handle, self.lib.hsc_b200_decode(
                self.handle, ctypes.c_void_p(p.data_ptr()), ctypes.c_void_p(i.data_ptr()), ctypes.c_void_p(c.data_ptr()),
                int(p.numel()), int(T), ctypes.c_void_p(out.data_ptr()), self._stream_ptr(stream)))
        return out


class EncodeSlot(object):
    """One encode in flight: a native view of the engine's dictionary with its own workspace (correlation map, argmax
    hierarchy, states), event buffers and residual.  begin() = K1 + state reset, run() = K2, both asynchronous on the
    stream given; states_async() copies the per-signal states to the slot's pinned buffer."""

    def __init__(self, eng, handle, S, T, capacity):
        torch = _torch()
        self.eng, self.handle = eng, handle
        self.S, self.T, self.cap = int(S), int(T), int(capacity)
        dev = eng.device
        with torch.cuda.device(dev):
            self.wsb = eng.workspace_bytes(S, T)
            self.ws = torch.empty((self.wsb,), dtype=torch.uint8, device=dev)
            self.resid = torch.empty((S, T, eng.F), dtype=eng.torch_dtype, device=dev)
            self.evp = torch.empty((S, capacity), dtype=torch.int32, device=dev)
            self.evi = torch.empty((S, capacity), dtype=torch.int32, device=dev)
            self.evc = torch.empty((S, capacity), dtype=eng.torch_dtype, device=dev)
            self.hstate = torch.empty((S * ctypes.sizeof(N.SignalState),), dtype=torch.uint8).pin_memory()
        self.states = (N.SignalState * S).from_address(self.hstate.data_ptr())
        self.k2_done = None

    def begin(self, xd, options, stream):
        e = self.eng
        N.check(e.lib, self.handle, e.lib.hsc_b200_mp_begin(
            self.handle, ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(self.resid.data_ptr()), self.S, self.T,
            ctypes.c_void_p(self.ws.data_ptr()), self.wsb, ctypes.byref(options), ctypes.c_void_p(stream.cuda_stream)))

    def run(self, stream):
        e = self.eng
        sp = ctypes.c_void_p(stream.cuda_stream)
        N.check(e.lib, self.handle, e.lib.hsc_b200_mp_run(
            self.handle, ctypes.c_void_p(self.evp.data_ptr()), ctypes.c_void_p(self.evi.data_ptr()),
            ctypes.c_void_p(self.evc.data_ptr()), self.cap, None, sp))
        N.check(e.lib, self.handle, e.lib.hsc_b200_mp_states_async(self.handle, self.states, sp))

    def compact(self, out, stream):
        e = self.eng
        torch = _torch()
        if out is None:
            dev = e.device
            out = dict(offsets=torch.empty((self.S + 1,), dtype=torch.int64, device=dev),
                       pos=torch.empty((self.S * self.cap,), dtype=torch.int32, device=dev),
                       idx=torch.empty((self.S * self.cap,), dtype=torch.int32, device=dev),
                       coef=torch.empty((self.S * self.cap,), dtype=e.torch_dtype, device=dev))
        N.check(e.lib, self.handle, e.lib.hsc_b200_mp_compact_events(
            self.handle, ctypes.c_void_p(self.evp.data_ptr()), ctypes.c_void_p(self.evi.data_ptr()), ctypes.c_void_p(self.evc.data_ptr()),
            self.cap, ctypes.c_void_p(out['offsets'].data_ptr()), ctypes.c_void_p(out['pos'].data_ptr()),
            ctypes.c_void_p(out['idx'].data_ptr()), ctypes.c_void_p(out['coef'].data_ptr()), self.S * self.cap,
            ctypes.c_void_p(stream.cuda_stream)))
        return out

    def launches(self):
        return int(self.eng.lib.hsc_b200_launch_count(self.handle))


_engines = {}
_dict_engines = {}


def engine_for_dictionary(D, weights=None, dtype=None, device=None, max_entries=16):
    """An engine that already holds dictionary D (weights, dtype): callers that alternate between a few dictionaries - the
    levels of a multilevel dictionary, their input-level representations for decoding - get one native engine per
    dictionary (D, Gram tensor and K1 operand stay on the device between calls) instead of re-uploading on every switch.
    Small LRU per (process, device), keyed by the dictionary's bytes."""
    import os
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError('hierarchical_sparse_coding_b200 needs a CUDA device (no CPU fallback)')
    idx = torch.cuda.current_device() if device is None else int(torch.device(device).index or 0) if not isinstance(device, int) else device
    D = np.asarray(D)
    dt = np.dtype(dtype) if dtype is not None else np.dtype(engine_dtype(D))
    Dc = np.ascontiguousarray(D if D.ndim == 3 else D[:, :, None], dtype=dt)
    wc = None if weights is None else np.ascontiguousarray(np.asarray(weights), dtype=dt)
    key = (os.getpid(), idx, str(dt), Dc.shape, hash(Dc.tobytes()), None if wc is None else hash(wc.tobytes()))
    cache = _dict_engines
    eng = cache.pop(key, None)
    if eng is None:
        eng = Engine(idx)
        eng.set_dictionary(Dc, weights=wc, dtype=dt)
        while len(cache) >= max_entries:
            old = cache.pop(next(iter(cache)))
            old.close()
    cache[key] = eng            # most recently used last
    return eng


def get_engine(device=None):
    """Process-wide engine per device (created lazily, after fork: the reference's Pool fan-out,
    scripts/scale_weight_effect_mlcsc.py:165, must not inherit a CUDA context)."""
    import os
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError('hierarchical_sparse_coding_b200 needs a CUDA device (no CPU fallback)')
    idx = torch.cuda.current_device() if device is None else int(torch.device(device).index or 0) if not isinstance(device, int) else device
    key = (os.getpid(), idx)
    if key not in _engines:
        _engines[key] = Engine(idx)
    return _engines[key]
