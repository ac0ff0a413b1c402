#!/bin/bash
# Grid / general / feature tests + own-arm bench of the general-path configurations + the bare-network training configurations with the reference.
O=gpurun_out/r2t; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_encoding.py tests/test_gpu_general.py tests/test_gpu_features.py tests/test_gpu_network.py -q -m gpu 2>&1 | tail -3
timeout 200 python scripts/bench_configs.py --steps 30 --configs image_w128,hash3d_w128,f4_l8,d4 > $O/own.jsonl 2> $O/err.log
timeout 400 python scripts/bench_configs.py --steps 30 --configs mlp_w128_h4,mlp_w64_h4 --reference > $O/mlp.jsonl 2>> $O/err.log
python - <<EOF2
import json
for f in ("own","mlp"):
    for l in open(f"$O/{f}.jsonl"):
        d=json.loads(l); r=d.get("reference",{})
        print(d["config"], "train %.4f inf %.4f"%(d["ms_per_step"],d["inference_ms"]), d.get("vs_reference_training"), d.get("vs_reference_inference"), {k:round(v.get("ms_per_step",-1),4) for k,v in r.items()})
EOF2
tail -3 $O/err.log
