mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; tail -n 2 gpurun_out/pytest_gpu.log
VARIANTS="1 4" bash tools/gpu_variants.sh
HSC_PURSUIT_VARIANT=4 HSC_B200_LIB=$PWD/hierarchical_sparse_coding_b200/libhsc_b200_prof.so timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_prof.log 2>&1
echo "512 signals v4:"; grep "hsc phases" gpurun_out/bench_c4_prof.log | tail -1
