#!/bin/bash
# Round-end dry run, as the driver does it: GPU tests, smoke(), both bench arms; plus the own arm of the secondary configurations.
O=gpurun_out/r2z; mkdir -p $O
timeout 700 python -m pytest tests -q -m gpu -x > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.log 2>&1; grep '^{' $O/bench_ref.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_own.log 2>&1; grep '^{' $O/bench_own.log | cut -c1-600
timeout 200 python scripts/bench_configs.py --steps 30 > $O/bench_configs_own.jsonl 2> $O/bench_configs.err
python - <<EOF2
import json
for l in open("$O/bench_configs_own.jsonl"):
    d=json.loads(l); print(d["config"], "train %.4f inf %.4f"%(d["ms_per_step"],d["inference_ms"]))
EOF2
