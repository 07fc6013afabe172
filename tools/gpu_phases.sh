mkdir -p gpurun_out
for pf in 0 1; do
HSC_PREFETCH=$pf HSC_B200_LIB=$PWD/hierarchical_sparse_coding_b200/libhsc_b200_prof.so timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_phases_pf$pf.log 2>&1
grep "hsc phases" gpurun_out/bench_phases_pf$pf.log | tail -2
done
HSC_K2_TMA=0 HSC_B200_LIB=$PWD/hierarchical_sparse_coding_b200/libhsc_b200_prof.so timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_phases_notma.log 2>&1
grep "hsc phases" gpurun_out/bench_phases_notma.log | tail -2
