"""A/B: the fused kernel against the general path (tcnnb_debug_set("general", 1)) on configurations both cover.

    python scripts/ab_paths.py [config,...] [batch]
"""
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

names = (sys.argv[1] if len(sys.argv) > 1 else "identity_cutlass,headline,image_w64").split(",")
for name in names:
    n_in, n_out, B, cfg, _ = CONFIGS[name]
    if len(sys.argv) > 2:
        B = int(sys.argv[2])
    for general in (0, 1, 0, 1):
        m = tcnn_b200.create_from_config(n_in, n_out, cfg)
        m.debug_set("general", general)
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
        t_train = e0.elapsed_time(e1) / 50
        e0.record()
        for i in range(50):
            m.network.inference(xs[i % 4])
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"config": name, "batch": B, "path": "general" if general else "fused", "train_ms": t_train, "inference_ms": e0.elapsed_time(e1) / 50, "loss": m.trainer.loss()}), flush=True)
