#!/bin/bash
# 8-GPU box: weak scaling at N=8 and N=4 (peer-memory engine; NCCL engine A/B at N=8), peer-memory parity at world 8.
O=gpurun_out/r2n8; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 400 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 30 --warmup 10 > $O/bench_n8.log 2>&1; echo "n8 rc=$?"
timeout 400 $TR --nproc-per-node 4 --master-port 29612 bench.py --gpus 4 --steps 30 --warmup 10 > $O/bench_n4.log 2>&1; echo "n4 rc=$?"
TCNNB_DP_NCCL=1 timeout 400 $TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --steps 30 --warmup 10 > $O/bench_n8_nccl.log 2>&1; echo "n8 nccl rc=$?"
TCNNB_DP_WORLD=8 TCNNB_DP_MODES=3,4 TCNNB_REQUIRE_PEER_MEMORY=1 timeout 400 python scripts/dp_parity.py > $O/dp_parity_n8.log 2>&1; echo "parity rc=$?"
timeout 200 python bench.py --steps 20 --warmup 5 > $O/bench_n1.log 2>&1; echo "n1 rc=$?"
grep -h '^{' $O/bench_n8.log $O/bench_n4.log $O/bench_n8_nccl.log $O/bench_n1.log | cut -c1-400
tail -3 $O/dp_parity_n8.log
