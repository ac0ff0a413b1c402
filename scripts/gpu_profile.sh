mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 6 > gpurun_out/launches.log 2>&1; echo "launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_step -s 3 -c 1 -f -o gpurun_out/prof_fused python scripts/profile_step.py 5 > gpurun_out/prof_fused.log 2>&1; echo "prof_fused rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:adam_step -s 3 -c 1 -f -o gpurun_out/prof_adam python scripts/profile_step.py 5 > gpurun_out/prof_adam.log 2>&1; echo "prof_adam rc=$?"
ls -la gpurun_out
