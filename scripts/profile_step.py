"""Runs a few training steps of the headline config (for ncu). Usage: python scripts/profile_step.py [steps] [batch]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-cuda-nn_b200"))
import torch

import tcnn_b200

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "configs", "headline.json")))
model = tcnn_b200.create_from_config(3, 3, cfg)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(B, 3, device="cuda", generator=g)
y = torch.rand(B, 3, device="cuda", generator=g)
for _ in range(steps):
    model.trainer.training_step(x, y)
torch.cuda.synchronize()
print("loss", model.trainer.loss())
