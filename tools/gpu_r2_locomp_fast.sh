# fast LoCOMP kernel (HSC_LOCOMP_FAST): A/B equality with the original kernel + reference traces, then timing on configs 4 / 5
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "locomp" -s > gpurun_out/pytest_gpu_locomp_fast.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|assert|events per signal|locomp" gpurun_out/pytest_gpu_locomp_fast.log | tail -25
for fast in 0 1 0 1; do
  echo "== HSC_LOCOMP_FAST=$fast"
  HSC_LOCOMP_FAST=$fast bash tools/gpu_locomp.sh 2>&1 | grep locomp
done
