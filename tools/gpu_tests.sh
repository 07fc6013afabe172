mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 30 gpurun_out/pytest_gpu.log
