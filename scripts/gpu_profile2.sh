mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_step -s 3 -c 1 -f -o gpurun_out/prof_fused_v2 python scripts/profile_step.py 5 > gpurun_out/prof_fused_v2.log 2>&1; echo "prof rc=$?"
TCNNB_ABLATE=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_step -s 3 -c 1 -f -o gpurun_out/prof_fused_v2_abl3 python scripts/profile_step.py 5 > gpurun_out/prof_fused_v2_abl3.log 2>&1; echo "prof abl rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_v2.csv python scripts/profile_step.py 6 > /dev/null 2>&1
ls -la gpurun_out | head -30
