"""End-to-end (host buffers) step time of the C ABI call, for A/B runs (TCNNB_NO_PDL=1, pageable vs pinned inputs)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-cuda-nn_b200"))
import torch

import tcnn_b200

B = 1 << 18
cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "configs", "headline.json")))
model = tcnn_b200.create_from_config(3, 3, cfg)
res = {"no_pdl": bool(os.environ.get("TCNNB_NO_PDL"))}
for kind in ("pinned", "pageable"):
    x, y = torch.rand(B, 3), torch.rand(B, 3)
    if kind == "pinned":
        x, y = x.pin_memory(), y.pin_memory()
    xn, yn = x.numpy(), y.numpy()
    for _ in range(5):
        model.training_step_host(xn, yn)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 50
    for _ in range(n):
        model.training_step_host(xn, yn)
    torch.cuda.synchronize()
    res[kind + "_ms"] = (time.perf_counter() - t0) / n * 1e3
xd, yd = torch.rand(B, 3, device="cuda"), torch.rand(B, 3, device="cuda")
for _ in range(5):
    model.trainer.training_step(xd, yd)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100):
    model.trainer.training_step(xd, yd)
torch.cuda.synchronize()
res["device_resident_ms"] = (time.perf_counter() - t0) / 100 * 1e3
print(json.dumps({"e2e_probe": res}))
