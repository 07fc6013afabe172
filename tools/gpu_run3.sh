mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c4.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c4.log
tail -n 3 gpurun_out/bench_c4.log
timeout 600 python bench.py --steps 1 --warmup 1 --signals 148 --no-cpu-baseline > gpurun_out/plain148.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --signals 148 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 600 python bench.py --steps 1 --warmup 1 --signals 148 --no-cpu-baseline > gpurun_out/plain148b.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:pursuit -s 1 -c 1 -o gpurun_out/prof_pursuit_r1 python bench.py --steps 1 --warmup 1 --signals 148 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
