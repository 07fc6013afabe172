mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
HSC_K2_L2HINTS=3 timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"pursuit" -s 1 -c 1 --csv --log-file gpurun_out/ncu_hints3.csv $CMD > gpurun_out/ncu_hints3.log 2>&1; echo "rc=$?"
grep -E "pursuit" gpurun_out/ncu_hints3.csv | cut -d, -f5,13- | head
