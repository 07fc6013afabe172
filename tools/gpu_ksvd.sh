mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k ksvd > gpurun_out/pytest_ksvd.log 2>&1; echo "pytest rc=$?"
tail -n 40 gpurun_out/pytest_ksvd.log
