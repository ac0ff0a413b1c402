mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 100 --warmup 10 > gpurun_out/bench_n4.log 2>&1; echo "n4 rc=$?"; tail -n 1 gpurun_out/bench_n4.log | cut -c1-1100
