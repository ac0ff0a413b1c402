mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_step -s 3 -c 1 -f -o gpurun_out/prof_fused_v3 python scripts/profile_step.py 5 > gpurun_out/prof_fused_v3.log 2>&1; echo "prof rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_v3.csv python scripts/profile_step.py 6 > /dev/null 2>&1
for a in 0 3; do TCNNB_ABLATE=$a timeout 120 python scripts/ablate.py; done 2>&1 | grep ablate | tee gpurun_out/ablate.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
