"""Phase timeline of the stand-alone MLP kernel (csrc/mlp_fused.cu) from its clock64 stamps: where a slot's layer period goes.

    python scripts/mlp_timeline.py [width] [hidden] > profiles/r02_mlp_timeline.json

Per slot and hidden layer the epilogue warp records: 0 wait start, 1 accumulator ready, then per chunk c (0 / 1): 2+3c tcgen05.ld
returned, 3+3c tcgen05.st issued, 4+3c stored + fenced + arrived on a_ready[s][c]; the MMA issuer records 0 first half's operand seen
ready, 1 first half issued, 2 second half issued + committed."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-cuda-nn_b200"))
import numpy as np
import torch

import tcnn_b200

width = int(sys.argv[1]) if len(sys.argv) > 1 else 128
hidden = int(sys.argv[2]) if len(sys.argv) > 2 else 4
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
slots = 2 if width == 128 else 4
n_ctas = 148
tiles_per_slot = 6
B = n_ctas * 128 * slots * tiles_per_slot
net = tcnn_b200.Network(width, width, {"otype": "FullyFusedMLP", "n_neurons": width, "n_hidden_layers": hidden})
tcnn_b200._check(tcnn_b200.load().tcnnb_network_debug_flags(net._h, flags))
p16 = net.initial_params(1).half().contiguous()
x = torch.rand(B, width, device="cuda").half().contiguous()
for _ in range(3):
    net.inference_mixed_precision(x, p16)
clocks = torch.zeros(n_ctas, 5, 64, 8, dtype=torch.int64, device="cuda")
tcnn_b200._check(tcnn_b200.load().tcnnb_network_debug_clocks(net._h, clocks.data_ptr()))
net.inference_mixed_precision(x, p16)
torch.cuda.synchronize()
tcnn_b200._check(tcnn_b200.load().tcnnb_network_debug_clocks(net._h, None))
c = clocks.cpu().numpy().astype(np.float64)
n_layers = hidden + 1
n_ev = min(64, tiles_per_slot * n_layers)
res = {"variant_flags": flags, "width": width, "n_hidden_layers": hidden, "slots": slots, "batch": B, "unit": "SM cycles (clock64), medians over CTAs / slots / steady-state events"}


def med(a):
    a = np.asarray(a)
    return float(np.median(a)) if a.size else None


ep = c[:, 1 : 1 + slots, :n_ev, :]  # [cta][slot][event][field]
hid = np.array([e for e in range(n_ev) if e % n_layers != hidden and e >= n_layers])  # hidden-layer events after the first tile
two = width >= 64
res["epilogue"] = {
    "wait_for_accumulator": med(ep[:, :, hid, 1] - ep[:, :, hid, 0]),
    "ld_chunk0": med(ep[:, :, hid, 2] - ep[:, :, hid, 1]),
    "convert_st_chunk0": med(ep[:, :, hid, 3] - ep[:, :, hid, 2]),
    "wait_st_fence_arrive_chunk0": med(ep[:, :, hid, 4] - ep[:, :, hid, 3]),
    "ld_chunk1": med(ep[:, :, hid, 5] - ep[:, :, hid, 4]) if two else None,
    "convert_st_chunk1": med(ep[:, :, hid, 6] - ep[:, :, hid, 5]) if two else None,
    "wait_st_fence_arrive_chunk1": med(ep[:, :, hid, 7] - ep[:, :, hid, 6]) if two else None,
    "busy_total": med(ep[:, :, hid, 7 if two else 4] - ep[:, :, hid, 1]),
    "layer_period": med(ep[:, :, hid[1:], 1] - ep[:, :, hid[1:] - 1, 1]) if len(hid) > 1 else None,
}
# whole tiles: accumulator-ready of layer 0 of consecutive tiles of a slot; output-layer epilogue + input staging = what is left
first = np.array([e for e in range(n_layers, n_ev - n_layers, n_layers)])
if len(first):
    res["tile"] = {"tile_period": med(ep[:, :, first + n_layers, 1] - ep[:, :, first, 1]) if first.max() + n_layers < n_ev else None,
                   "last_hidden_ready_to_next_tile_layer0_ready (output epilogue + input staging + layer-0 MMA)": med(ep[:, :, first[1:], 1] - ep[:, :, first[1:] - 1, 1]) if len(first) > 1 else None}
iss = c[:, 0, : min(64, n_ev * slots), :]
ev = np.arange(slots * n_layers, iss.shape[1])
res["issuer"] = {
    "issue_first_half": med(iss[:, ev, 1] - iss[:, ev, 0]),
    "first_half_issued_to_second_half_issued_and_committed": med(iss[:, ev, 2] - iss[:, ev, 1]) if two else None,
}
lat_acc = []
for s in range(slots):
    for e in hid:
        cur = e * slots + s
        if cur < iss.shape[1]:
            lat_acc.append(ep[:, s, e, 1] - iss[:, cur, 2 if two else 1])
res["handoff"] = {"commit_to_accumulator_seen (MMA drain + barrier)": med(np.concatenate(lat_acc)) if lat_acc else None}
total = c[:, 1 : 1 + slots, :n_ev, 1]
res["mma_floor_cycles_per_layer"] = 128 * width * width / 8192 / 4 if False else (width * width) / 32
# raw timeline of one CTA (cycles relative to its first stamp): issuer events [slot][layer-event][fields], epilogue events per slot
cta = c[3]
t0 = np.nanmin(np.where(cta > 0, cta, np.nan))
raw = {"issuer": [[int(v - t0) if v > 0 else None for v in cta[0, e, :3]] for e in range(min(40, iss.shape[1]))]}
for s_ in range(slots):
    raw[f"slot{s_}"] = [[int(v - t0) if v > 0 else None for v in cta[1 + s_, e, :8]] for e in range(min(20, n_ev))]
res["raw_cta3"] = raw
print(json.dumps(res))
