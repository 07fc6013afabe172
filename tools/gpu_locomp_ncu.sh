# ncu --set full of the LoCOMP kernel on the config-4 shard (second launch), after a plain run has exited 0
mkdir -p gpurun_out
timeout 300 python tools/locomp_c4.py > gpurun_out/locomp_c4.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/locomp_c4.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"locomp" -s 1 -c 1 -f -o gpurun_out/prof_r2_locomp python tools/locomp_c4.py > gpurun_out/ncu_full_locomp.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_r2_locomp.ncu-rep > gpurun_out/prof_r2_locomp.json 2>/dev/null
