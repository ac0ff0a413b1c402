"""A/B of the dense-level replicas inside the fused kernel (tcnnb_debug_set("fused_replicas", n)) on the headline configuration."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-cuda-nn_b200"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import tcnn_b200  # noqa: E402
from bench import closed_form_targets  # noqa: E402
from bench_configs import CONFIGS  # noqa: E402

n_in, n_out, B, cfg, _ = CONFIGS["headline"]
for r in (0, 8, 32, 0, 16):
    m = tcnn_b200.create_from_config(n_in, n_out, cfg)
    m.debug_set("fused_replicas", r)
    xs = [torch.rand(B, n_in, device="cuda") for _ in range(4)]
    ys = [closed_form_targets(torch, x, n_out) for x in xs]
    for i in range(10):
        m.trainer.training_step(xs[i % 4], ys[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        m.trainer.training_step(xs[i % 4], ys[i % 4])
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"fused_replicas": r, "ms_per_step": e0.elapsed_time(e1) / 50, "loss": m.trainer.loss()}), flush=True)
