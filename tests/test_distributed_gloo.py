"""The N>1 host logic on CPU: world_size-2 gloo processes shard the signals, each 'encodes' its
shard (events are synthesised here - no GPU), rank 0 gathers the codes; the union must equal the
single-process answer in global signal order."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hierarchical_sparse_coding_b200 import distributed as hd   # noqa: E402


def _fake_events(signal_index):
    rs = np.random.RandomState(100 + signal_index)
    n = int(rs.randint(0, 9))
    return rs.randint(0, 1000, n).astype(np.int32), rs.randint(0, 16, n).astype(np.int32), rs.randn(n)


def _worker(rank, world, port, n_signals, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = hd.shard_range(n_signals, rank, world)
    ev = [_fake_events(i) for i in range(lo, hi)]
    counts, pos, idx, coef = hd.pack_events([e[0] for e in ev], [e[1] for e in ev], [e[2] for e in ev])
    out = hd.gather_events(counts, pos, idx, coef, dst=0)
    if rank == 0:
        allp, alli, allc = [], [], []
        for (c, p, i, v) in out:
            pp, ii, cc = hd.unpack_events(c, p, i, v)
            allp += pp
            alli += ii
            allc += cc
        q.put([(p.tolist(), i.tolist(), c.tolist()) for p, i, c in zip(allp, alli, allc)])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 512, 4096, 4099):
        for w in (1, 2, 3, 8):
            spans = [hd.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    ev = [_fake_events(i) for i in range(11)]
    c, p, i, v = hd.pack_events([e[0] for e in ev], [e[1] for e in ev], [e[2] for e in ev])
    pp, ii, vv = hd.unpack_events(c, p, i, v)
    for a, b in zip(ev, zip(pp, ii, vv)):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_two_rank_gather_matches_single_process():
    n_signals = 7
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_signals, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert len(got) == n_signals
    for i, (p, k, c) in enumerate(got):
        ep, ek, ec = _fake_events(i)
        assert p == ep.tolist() and k == ek.tolist() and np.allclose(c, ec)


def _worker_device_gather(rank, world, port, n_signals, q):
    """gather_device_events on compacted codes (CPU tensors under gloo; device tensors under NCCL on the GPU box)."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = hd.shard_range(n_signals, rank, world)
    ev = [_fake_events(i) for i in range(lo, hi)]
    counts = np.array([len(e[0]) for e in ev], dtype=np.int64)
    cat = lambda j, dt: np.concatenate([e[j] for e in ev]).astype(dt) if counts.sum() else np.zeros(0, dt)    # noqa: E731
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    cap = 64                                                 # the compacted buffers are longer than the code, like the engine's
    pad = lambda a: torch.from_numpy(np.concatenate([a, np.zeros(cap - len(a), a.dtype)]))                       # noqa: E731
    dev = dict(offsets=torch.from_numpy(offsets), pos=pad(cat(0, np.int32)), idx=pad(cat(1, np.int32)), coef=pad(cat(2, np.float32)),
               total=int(counts.sum()))
    out = hd.gather_device_events(dev, dst=0, host_out=True)
    if rank == 0:
        res = []
        for r in range(world):
            off = out['offsets'][r]
            for s in range(len(off) - 1):
                a, b = int(off[s]), int(off[s + 1])
                res.append((out['pos'][r][a:b].tolist(), out['idx'][r][a:b].tolist(), out['coef'][r][a:b].tolist()))
        q.put(res)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_device_gather_of_compacted_codes_world2():
    """The gather of the sparse codes as the GPU path does it (offsets all-gathered, arrays padded to the largest rank):
    the union over the ranks equals the single-process answer in global signal order."""
    n_signals = 8          # equal shards: the offsets vectors all-gather as equal-sized tensors
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_device_gather, args=(r, 2, port, n_signals, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(got) == n_signals
    for i, (p, k, c) in enumerate(got):
        rp, rk, rc = _fake_events(i)
        assert p == rp.tolist() and k == rk.tolist() and np.allclose(c, rc.astype(np.float32))
