mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_ws -s 3 -c 1 -f -o gpurun_out/prof_fused_ws_final python scripts/profile_step.py 5 > gpurun_out/prof_fused_ws_final.log 2>&1; echo "prof rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_final.csv python scripts/profile_step.py 6 > /dev/null 2>&1; echo "launches rc=$?"
