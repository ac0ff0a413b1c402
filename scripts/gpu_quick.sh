mkdir -p gpurun_out
timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee gpurun_out/ablate.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_own.log 2>&1; echo "own rc=$?"; tail -n 1 gpurun_out/bench_own.log | cut -c1-400
