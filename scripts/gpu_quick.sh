mkdir -p gpurun_out
export ABL=$PWD/scripts/experiments/libtcnn_b200_ablation.so
export NOQ=$PWD/scripts/experiments/libtcnn_b200_noquad.so
TCNNB_LIB=$NOQ timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee gpurun_out/ablate_noquad.log
timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee gpurun_out/ablate.log
TCNNB_LIB=$ABL timeout 200 python scripts/ws_timeline.py > gpurun_out/ws_timeline.log 2>&1; echo "timeline rc=$?"; head -n 1 gpurun_out/ws_timeline.log | cut -c1-1500
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_own.log 2>&1; tail -n 1 gpurun_out/bench_own.log | cut -c1-400
