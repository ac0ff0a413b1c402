mkdir -p gpurun_out
timeout 300 python tests/debug_stages.py hash3d_small 3 3 512 > gpurun_out/debug_ws.log 2>&1; echo "debug rc=$?"; grep -E "encoded|hidden\[|output |dW\[|grid grads|params after|loss dev" gpurun_out/debug_ws.log | cut -c1-200
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee gpurun_out/ablate.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_own.log 2>&1; tail -n 1 gpurun_out/bench_own.log | cut -c1-300
