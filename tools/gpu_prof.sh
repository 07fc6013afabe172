mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --signals 148 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain148.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
timeout 600 $CMD > gpurun_out/plain148b.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"correlate_tc|pursuit" -s 2 -c 2 -o gpurun_out/prof_k1k2 $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
