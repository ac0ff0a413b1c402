"""Phase ablation timing of the fused kernel (profiling experiment). Run under TCNNB_ABLATE=<bits> (1 = no gather loads,
2 = no scatter reductions, 4 = no 64-bit pairing); prints the device time of the fused kernel and of Adam."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-cuda-nn_b200"))
import torch

import tcnn_b200

B = 1 << 18
cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "configs", "headline.json")))
model = tcnn_b200.create_from_config(3, 3, cfg)
g = torch.Generator(device="cuda").manual_seed(1)
xs = [torch.rand(B, 3, device="cuda", generator=g) for _ in range(4)]
ys = [torch.rand(B, 3, device="cuda", generator=g) for _ in range(4)]
for i in range(10):
    model.trainer.training_step(xs[i % 4], ys[i % 4])
torch.cuda.synchronize()
model.set_profiling(True)
N = 50
for i in range(N):
    model.trainer.training_step(xs[i % 4], ys[i % 4])
p = model.read_profile()
print(json.dumps({"ablate": int(os.environ.get("TCNNB_ABLATE", "0")), "fused_ms": p["fused_ms_total"] / p["n_steps"], "binning_ms": p["binning_ms_total"] / p["n_steps"], "adam_ms": p["optimizer_ms_total"] / p["n_steps"]}))
