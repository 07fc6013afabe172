"""Worker of tests/test_parity_gpu.py::test_sharded_encode_equals_single_gpu (N ranks, one GPU each): the signals of a batch
are partitioned over the ranks by contiguous index ranges (distributed.shard_range), every rank encodes its
shard through the public host pipeline, the sparse codes are gathered to rank 0 from the device buffers
(distributed.gather_device_events, NCCL) - and must equal, signal for signal and bit for bit, what one GPU returns for the
whole batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hierarchical_sparse_coding_b200 as hsc          # noqa: E402
from hierarchical_sparse_coding_b200 import distributed as hd   # noqa: E402


def main():
    rank = int(os.environ['RANK'])
    world = int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rs = np.random.RandomState(23)
    S, T, F, K, L, n = 4 * world + 2, 8192, 4, 64, 32, 60          # ragged: the shards differ in size
    D = rs.randn(K, L, F)
    D = (D / np.sqrt(np.sum(D * D, axis=(1, 2), keepdims=True))).astype(np.float32)
    x = np.zeros((S, T, F), np.float32)
    for s in range(S):
        for p, k, a in zip(rs.randint(0, T - L, n), rs.randint(0, K, n), rs.uniform(0.25, 4.0, n)):
            x[s, p:p + L] += np.float32(a) * D[k]
    eng = hsc.Engine(local)
    eng.set_dictionary(D)
    opt = eng.make_options(nbNonzeroCoefs=n)
    lo, hi = hd.shard_range(S, rank, world)
    batches = [torch.from_numpy(x[lo:hi]).pin_memory()] * 3         # three identical batches through the pipeline
    gathered = []
    for r in eng.encode_host_pipelined(batches, opt, capacity=256, n_chunks=2, host_events=False,
                                       on_device_events=lambda bi, dev: hd.gather_device_events(dev, dst=0, host_out=True)):
        gathered.append(r.gathered)
        assert r.total_events() >= n * (hi - lo)
    if rank == 0:
        ref = eng.encode_host(torch.from_numpy(x).pin_memory(), opt, capacity=256)
        for g in gathered:
            s = 0
            for rr in range(world):
                off = g['offsets'][rr]
                assert len(off) - 1 == hd.shard_range(S, rr, world)[1] - hd.shard_range(S, rr, world)[0]
                for j in range(len(off) - 1):
                    a, b = int(off[j]), int(off[j + 1])
                    assert np.array_equal(g['pos'][rr][a:b], ref.pos[s]) and np.array_equal(g['idx'][rr][a:b], ref.idx[s]), (rr, j)
                    assert np.array_equal(g['coef'][rr][a:b], ref.coef[s]), (rr, j)
                    s += 1
            assert s == S
        print('sharded encode over %d GPUs == single GPU: %d signals, %d atoms, bit for bit' % (world, S, ref.total_events()))
    else:
        assert all(g is None for g in gathered)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
