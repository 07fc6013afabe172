# ncu source-level capture of K2 on the single-sequence shapes (config 2 shape with 3000 atoms, config 1): where the per-selection
# latency chain stalls.  Reports come back under gpurun_out/ (read with ncu -i ... --page source --csv).
mkdir -p gpurun_out
cat > /tmp/single_run.py <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, '.')
import bench
import hierarchical_sparse_coding_b200 as hsc
name = sys.argv[1]
w = dict(bench.WORKLOADS['c2'], atoms=3000) if name == 'c2' else dict(S=1, T=10000, F=1, K=4, L=16, atoms=289)
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
    eng.run_only(evp, evi, evc, cap, sync_states=True)
torch.cuda.synchronize()
PY
for n in c2 c1; do
  timeout 600 python /tmp/single_run.py $n && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:pursuit_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_single_$n python /tmp/single_run.py $n > gpurun_out/ncu_single_$n.log 2>&1
  echo "$n rc=$?"
done
ls -la gpurun_out/prof_r2_single_*
