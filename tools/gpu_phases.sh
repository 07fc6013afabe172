mkdir -p gpurun_out
HSC_PROF_DUMP=$PWD/gpurun_out/prof_dump.csv HSC_B200_LIB=$PWD/hierarchical_sparse_coding_b200/libhsc_b200_prof.so timeout 600 python - > gpurun_out/prof_dump.log 2>&1 <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, '.')
import bench
import hierarchical_sparse_coding_b200 as hsc
w = dict(bench.WORKLOADS['c4'])
D = bench.make_dictionary(w)
x = bench.make_signals(w, D, seed=1000)
eng = hsc.Engine(0); eng.set_dictionary(D)
opt = eng.make_options(nbNonzeroCoefs=w['atoms'])
xd = torch.from_numpy(x).cuda()
cap = w['atoms'] * 4 + 64
for _ in range(2):
    evp, evi, evc, states, resid = eng.encode_device(xd, opt, cap)
off = bench.centre_offset(w['L'])
L, T = w['L'], w['T']
pos = evp.cpu().numpy()
ne = np.array([st.n_events for st in states])
edges = []
for s in range(len(states)):
    t = pos[s, :ne[s]]
    edges.append(int(np.sum((t - (L - 1) < off) | (t + (L - 1) > T - L + off))))
np.savetxt('gpurun_out/prof_signals.csv', np.stack([ne, np.array(edges)], 1), fmt='%d', delimiter=',', header='n_events,n_edge')
PY
tail -3 gpurun_out/prof_dump.log
