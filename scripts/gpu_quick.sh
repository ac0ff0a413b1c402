mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_own.log 2>&1; echo "own rc=$?"; tail -n 1 gpurun_out/bench_own.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}); print(d['roofline']['kernel_ms'], d['roofline']['frac'])"
python -c "from __graft_entry__ import smoke; smoke()" 2>&1 | tail -n 1
