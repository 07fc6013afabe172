# per-phase cycle counts of K2 (HSC_PROFILE_PHASES build) with the near-tie re-ranking on / off / watch only
mkdir -p gpurun_out
for rr in 0 1e-12 4e-6; do
echo "== HSC_RERANK=$rr"
HSC_RERANK=$rr HSC_B200_LIB=$PWD/hierarchical_sparse_coding_b200/libhsc_b200_prof.so timeout 600 python - 2>&1 <<'PY' | grep -E "hsc phases|hsc timeline|hsc edge|K2 ms"
import numpy as np, torch, sys, time
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
for rep in range(3):
    resid = torch.empty_like(xd)
    evp = torch.empty((512, cap), dtype=torch.int32, device='cuda'); evi = torch.empty_like(evp); evc = torch.empty((512, cap), dtype=torch.float32, device='cuda')
    eng.begin_only(xd, opt, resid)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run_only(evp, evi, evc, cap, sync_states=True); e1.record(); torch.cuda.synchronize()
    print('K2 ms %.2f' % e0.elapsed_time(e1))
PY
done
