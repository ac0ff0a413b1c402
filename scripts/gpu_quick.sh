mkdir -p gpurun_out
timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee gpurun_out/ablate.log
timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee -a gpurun_out/ablate.log
timeout 200 ncu --metrics launch__shared_mem_config_size,gpu__time_duration.sum --clock-control none -k regex:fused_ws -s 3 -c 1 python scripts/profile_step.py 5 2>&1 | grep -E "shared_mem_config|time_duration" | tee gpurun_out/carveout.log
timeout 300 python -m pytest tests -x -q -m gpu -k "golden or stagewise or full_size or module" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_gpu.log
