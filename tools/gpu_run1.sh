set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload tiny --steps 2 --warmup 1 > gpurun_out/bench_tiny.log 2>&1; echo "rc=$?" >> gpurun_out/bench_tiny.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c4.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c4.log
tail -5 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench_tiny.log gpurun_out/bench_c4.log
