"""A/B timing of the two inference kernels (fused_ws.cu TRAIN=false vs fused_step.cu) on the headline configuration."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-cuda-nn_b200"))
import torch

import tcnn_b200

cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "configs", "headline.json")))
B = 1 << 18
m = tcnn_b200.create_from_config(3, 3, cfg)
rng = tcnn_b200.Pcg32(1337)
xs = [tcnn_b200.generate_random_uniform(rng, B * 3).view(B, 3) for _ in range(4)]
res = {}
for name, flag in (("ws", 0), ("sync", 1), ("ws", 0), ("sync", 1)):
    m.debug_set("inference_sync_kernel", flag)
    outs = []
    for i in range(5):
        outs.append(m.network.inference(xs[i % 4]))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        m.network.inference(xs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    res.setdefault(name, []).append(e0.elapsed_time(e1) / 50)
    res[name + "_out"] = outs[0].clone()
print(json.dumps({"ws_ms": res["ws"], "sync_ms": res["sync"], "identical": bool(torch.equal(res["ws_out"], res["sync_out"]))}))
