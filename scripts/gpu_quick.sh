mkdir -p gpurun_out
timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee gpurun_out/ablate.log
timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee -a gpurun_out/ablate.log
timeout 300 python -m pytest tests -x -q -m gpu -k "golden or stagewise or smoke or full_size" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_gpu.log
