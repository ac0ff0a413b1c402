#!/bin/bash
# Whole GPU suite + secondary-configuration bench (with the reference) + ncu captures of the general path's kernels.
O=gpurun_out/r2r; mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 700 python scripts/bench_configs.py --steps 30 --reference > $O/bench_configs.jsonl 2> $O/bench_configs.err
python - <<EOF2
import json
for l in open("$O/bench_configs.jsonl"):
    d=json.loads(l); print(d["config"], round(d["ms_per_step"],4), "%.3e"%d["samples_per_s"], d.get("vs_reference_training"), d.get("vs_reference_inference"))
EOF2
bash scripts/gpu_profile_general.sh 2>&1 | tail -6
