mkdir -p gpurun_out
timeout 400 python scripts/dp_parity.py > gpurun_out/dp_parity.log 2>&1; echo "dp_parity rc=$?"; grep -E "sharded|identical|max|OK|Error|error" gpurun_out/dp_parity.log | tail -n 16
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dp_profile.py > gpurun_out/dp_profile.log 2>&1; echo "rc=$?"; tail -n 1 gpurun_out/dp_profile.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2.log 2>&1; echo "n2 rc=$?"; tail -n 1 gpurun_out/bench_n2.log | cut -c1-700
