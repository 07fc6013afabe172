mkdir -p gpurun_out
for v in 1 0; do
HSC_PURSUIT_VARIANT=$v HSC_B200_LIB=$PWD/hierarchical_sparse_coding_b200/libhsc_b200_prof.so timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --signals 148 > gpurun_out/bench_c4_prof148.log 2>&1
echo "variant $v 148 signals:"; grep "hsc phases" gpurun_out/bench_c4_prof148.log | tail -1
done
HSC_B200_LIB=$PWD/hierarchical_sparse_coding_b200/libhsc_b200_prof.so timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_prof.log 2>&1
echo "512 signals:"; grep "hsc phases" gpurun_out/bench_c4_prof.log | tail -1
