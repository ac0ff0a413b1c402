mkdir -p gpurun_out
for a in 0 1 2 3; do TCNNB_ABLATE=$a timeout 120 python scripts/ablate.py; done 2>&1 | grep ablate | tee gpurun_out/ablate.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log
