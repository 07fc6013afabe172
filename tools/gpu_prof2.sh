mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
export HSC_PURSUIT_VARIANT=4
timeout 600 $CMD > gpurun_out/plain512.log 2>&1 && \
timeout 2400 ncu --set full --clock-control none --import-source on -k regex:"pursuit" -s 1 -c 1 -o gpurun_out/prof_k2_512 $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
