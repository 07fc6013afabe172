mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c4.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c4.log
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/bench_c2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c2.log
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/bench_c4.log gpurun_out/bench_c2.log
