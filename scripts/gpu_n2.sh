mkdir -p gpurun_out
timeout 300 python scripts/dp_parity.py > gpurun_out/dp_parity.log 2>&1; echo "dp_parity rc=$?"; tail -n 12 gpurun_out/dp_parity.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2.log 2>&1; echo "n2 rc=$?"; tail -n 1 gpurun_out/bench_n2.log | cut -c1-900
TCNNB_DP_REPLICATED=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2_repl.log 2>&1; echo "n2 repl rc=$?"; tail -n 1 gpurun_out/bench_n2_repl.log | cut -c1-400
