mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_gpu.log
for a in 0 1 2 3; do TCNNB_ABLATE=$a timeout 120 python scripts/ablate.py; done 2>&1 | grep ablate | tee gpurun_out/ablate.log
TCNNB_BINNING=0 timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee -a gpurun_out/ablate.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_own.log 2>&1; tail -n 1 gpurun_out/bench_own.log | cut -c1-300
