mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee gpurun_out/ablate.log
timeout 600 python bench.py > gpurun_out/bench_own.log 2>&1; echo "own rc=$?"; tail -n 1 gpurun_out/bench_own.log | cut -c1-400
compute-sanitizer --tool racecheck --print-limit 5 python scripts/profile_step.py 2 16384 > gpurun_out/racecheck.log 2>&1; tail -n 4 gpurun_out/racecheck.log
