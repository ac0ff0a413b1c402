mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_gpu.log
for a in 0 3; do TCNNB_ABLATE=$a timeout 120 python scripts/ablate.py; done 2>&1 | grep ablate | tee gpurun_out/ablate.log
TCNNB_BINNING=0 timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee -a gpurun_out/ablate.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_v4.csv python scripts/profile_step.py 6 > /dev/null 2>&1
grep -E "bin_|fused|adam" gpurun_out/launches_v4.csv | awk -F'","' '{print $5, $NF}' | sed 's/(.*)//' | tail -5
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_own.log 2>&1; tail -n 1 gpurun_out/bench_own.log | cut -c1-300
