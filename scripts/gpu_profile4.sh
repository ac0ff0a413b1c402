mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_ws -s 3 -c 1 -f -o gpurun_out/prof_fused_ws python scripts/profile_step.py 5 > gpurun_out/prof_fused_ws.log 2>&1; echo "prof rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ws.csv python scripts/profile_step.py 6 > /dev/null 2>&1
grep -E "bin_|fused|adam" gpurun_out/launches_ws.csv | awk -F'","' '{print $5, $NF}' | sed 's/(.*)//' | tail -5
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_own.log 2>&1; tail -n 1 gpurun_out/bench_own.log | cut -c1-300
