#!/bin/bash
# A/B runs of kernel variants on one box: N-split MLP issue (variant_nsplit build), dense-level replicas in the fused kernel.
O=gpurun_out/r2s; mkdir -p $O
V=tiny-cuda-nn_b200/variant_nsplit/libtcnn_b200.so
TCNNB_LIB=$V timeout 300 python -m pytest tests/test_gpu_network.py tests/test_gpu_general.py -q -m gpu 2>&1 | tail -3
for lib in "" $V; do
  echo "lib=${lib:-production}"
  TCNNB_LIB=$lib timeout 150 python scripts/bench_mlp.py --widths 128 --hidden 2,4,8 --batches 1048576 2>&1 | grep '^{' | cut -c1-330
done | tee $O/nsplit.txt
timeout 100 python scripts/ab_fused_replicas.py | tee $O/fused_replicas.txt
