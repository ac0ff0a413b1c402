mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2.log 2>&1; echo "n2 rc=$?"; tail -n 2 gpurun_out/bench_n2.log | cut -c1-600
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 10 > gpurun_out/bench_n1.log 2>&1; tail -n 1 gpurun_out/bench_n1.log | cut -c1-300
timeout 300 python scripts/dp_parity.py > gpurun_out/dp_parity.log 2>&1; echo "dp_parity rc=$?"; tail -n 5 gpurun_out/dp_parity.log
