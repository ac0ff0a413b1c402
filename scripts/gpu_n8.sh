mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/smi_n8.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/bench_n8.log 2>&1; echo "n8 rc=$?"; tail -n 1 gpurun_out/bench_n8.log | cut -c1-700
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 100 --warmup 10 > gpurun_out/bench_n4.log 2>&1; echo "n4 rc=$?"; tail -n 1 gpurun_out/bench_n4.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 scripts/dp_profile.py > gpurun_out/dp_profile_n8.log 2>&1; echo "prof rc=$?"; tail -n 1 gpurun_out/dp_profile_n8.log | cut -c1-900
