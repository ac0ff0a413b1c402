mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
bash tests/golden/make_golden.sh gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden rc=$?"
timeout 150 python tests/debug_stages.py hash3d_small 3 3 512 > gpurun_out/debug_hash3d.log 2>&1; echo "debug1 rc=$?"
timeout 150 python tests/debug_stages.py dense_mix3d 3 2 256 > gpurun_out/debug_dense.log 2>&1; echo "debug2 rc=$?"
timeout 150 python tests/debug_stages.py image2d 2 3 512 > gpurun_out/debug_image2d.log 2>&1; echo "debug3 rc=$?"
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py --impl reference --steps 100 --warmup 20 > gpurun_out/bench_ref.log 2>&1; echo "bench_ref rc=$?"
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_own.log 2>&1; echo "bench_own rc=$?"
tail -n 30 gpurun_out/debug_hash3d.log
tail -n 5 gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/bench_ref.log gpurun_out/bench_own.log
