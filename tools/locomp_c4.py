"""LoCOMP on the config-4 shard (512 signals x 65536 x 4, 256 filters x 64, 655 atoms each): two encodes, prints the
second one's time.  Used under ncu to capture the LoCOMP kernel (tools/gpu_locomp_ncu.sh)."""
import sys
import time

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
import hierarchical_sparse_coding_b200 as hsc  # noqa: E402

w = dict(bench.WORKLOADS['c4'])
w['S'] = int(sys.argv[1]) if len(sys.argv) > 1 else 512
D = bench.make_dictionary(w)
x = bench.make_signals(w, D, seed=1000)
eng = hsc.Engine(0)
eng.set_dictionary(D)
xd = torch.from_numpy(x).cuda()
opt = eng.make_options(nbNonzeroCoefs=w['atoms'], method=1)
for it in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    evp, evi, evc, states, resid = eng.encode_device(xd, opt, w['atoms'] * 8 + 256)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
nnz = sum(s.nnz for s in states)
print('locomp c4 S=%d: %.1f ms, %.3g atoms/s' % (w['S'], 1e3 * dt, nnz / dt))
