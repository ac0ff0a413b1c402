#!/bin/bash
# ncu --set full of the general path's network kernels (dgrad chain, weight gradients) and the replicated grid scatter, configs[2].
O=gpurun_out/prof_general; mkdir -p $O
for k in "mlp_forward_kernel<128u, false, true>:dgrad" "mlp_wgrad_kernel:wgrad" "grid_backward_kernel:gridbwd"; do
  name=${k%%:*}; tag=${k##*:}
  timeout 280 ncu --set full --clock-control none --import-source on -k regex:"${name%%<*}" -c 3 -o $O/prof_$tag -f python scripts/bench_configs.py --configs image_w128 --steps 1 --warmup 1 > $O/prof_$tag.log 2>&1
done
ls -la $O
