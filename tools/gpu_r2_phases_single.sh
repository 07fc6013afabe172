# per-phase cycle counts of K2 on the single-sequence shapes (config 1 and config 2), HSC_PROFILE_PHASES build
mkdir -p gpurun_out
HSC_K1=simt HSC_B200_LIB=$PWD/hierarchical_sparse_coding_b200/libhsc_b200_prof.so timeout 600 python - 2>&1 <<'PY'
import numpy as np, torch, sys, time
sys.path.insert(0, '.')
import bench
import hierarchical_sparse_coding_b200 as hsc
for name, w in (('c2', dict(bench.WORKLOADS['c2'], atoms=3000)), ('c1', dict(S=1, T=10000, F=1, K=4, L=16, atoms=289))):
    print('==', name, flush=True)
    D = bench.make_dictionary(w)
    x = bench.make_signals(w, D, seed=1000)
    eng = hsc.Engine(0); eng.set_dictionary(D)
    opt = eng.make_options(nbNonzeroCoefs=w['atoms'])
    xd = torch.from_numpy(x).cuda()
    cap = w['atoms'] * 4 + 64
    for rep in range(2):
        resid = torch.empty_like(xd)
        evp = torch.empty((1, cap), dtype=torch.int32, device='cuda'); evi = torch.empty_like(evp); evc = torch.empty((1, cap), dtype=torch.float32, device='cuda')
        eng.begin_only(xd, opt, resid)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run_only(evp, evi, evc, cap, sync_states=True); e1.record(); torch.cuda.synchronize()
        sys.stderr.flush()
        print('K2 ms %.2f  (%d atoms)' % (e0.elapsed_time(e1), w['atoms']), flush=True)
PY
