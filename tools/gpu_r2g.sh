# LoCOMP golden at the tightened tolerances, config 3 (incl. locomp), host profile of the hierarchical encode
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -rA -k "config3 or locomp or dropin" > gpurun_out/pytest_gpu_r2g.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|nnz|Error|events, engine|step-identical" gpurun_out/pytest_gpu_r2g.log | tail -n 40
timeout 300 python tools/prof_c3_host.py > gpurun_out/prof_c3_r2g.log 2>&1; head -n 60 gpurun_out/prof_c3_r2g.log
timeout 300 python tools/latency_c1_c3.py > gpurun_out/latency_r2g.log 2>&1; tail -n 4 gpurun_out/latency_r2g.log
