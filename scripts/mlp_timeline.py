"""Phase timeline of the stand-alone MLP kernel (csrc/mlp_fused.cu) from its clock64 stamps: where a slot's layer period goes.

    python scripts/mlp_timeline.py [width] [hidden] > profiles/r02_mlp_timeline.json

Per slot and hidden layer the epilogue warp (group 0, warp 0) records: 0 wait start, 1 accumulator ready, 2 tcgen05.ld of all of its
columns returned (before the slot's named barrier), 3 converted + tcgen05.st issued, 4 stored + fenced + arrived on a_ready[s];
the MMA issuer records 0 operand seen ready, 1 MMAs of the layer issued + committed."""
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
slots = 2 if width == 128 else 4
n_ctas = 148
tiles_per_slot = 6
B = n_ctas * 128 * slots * tiles_per_slot
net = tcnn_b200.Network(width, width, {"otype": "FullyFusedMLP", "n_neurons": width, "n_hidden_layers": hidden})
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
res = {"width": width, "n_hidden_layers": hidden, "slots": slots, "batch": B, "unit": "SM cycles (clock64), medians over CTAs / slots / steady-state events"}


def med(a):
    a = np.asarray(a)
    return float(np.median(a)) if a.size else None


ep = c[:, 1 : 1 + slots, :n_ev, :]  # [cta][slot][event][field]
hid = np.array([e for e in range(n_ev) if e % n_layers != hidden and e >= n_layers])  # hidden-layer events after the first tile
res["epilogue"] = {
    "wait_for_accumulator": med(ep[:, :, hid, 1] - ep[:, :, hid, 0]),
    "ld_all_columns": med(ep[:, :, hid, 2] - ep[:, :, hid, 1]),
    "group_barrier_convert_st": med(ep[:, :, hid, 3] - ep[:, :, hid, 2]),
    "wait_st_fence_arrive": med(ep[:, :, hid, 4] - ep[:, :, hid, 3]),
    "busy_total": med(ep[:, :, hid, 4] - ep[:, :, hid, 1]),
    "layer_period": med(ep[:, :, hid[1:], 1] - ep[:, :, hid[1:] - 1, 1]) if len(hid) > 1 else None,
}
# whole tiles: accumulator-ready of layer 0 of consecutive tiles of a slot; output-layer epilogue + input staging = what is left
first = np.array([e for e in range(n_layers, n_ev - n_layers, n_layers)])
if len(first):
    res["tile"] = {"tile_period": med(ep[:, :, first + n_layers, 1] - ep[:, :, first, 1]) if first.max() + n_layers < n_ev else None,
                   "last_hidden_ready_to_next_tile_layer0_ready (last MMA + output epilogue + input staging + layer-0 MMA)": med(ep[:, :, first[1:], 1] - ep[:, :, first[1:] - 1, 1]) if len(first) > 1 else None}
iss = c[:, 0, : min(64, n_ev * slots), :]
ev = np.arange(slots * n_layers, iss.shape[1])
res["issuer"] = {"issue_layer_and_commit": med(iss[:, ev, 1] - iss[:, ev, 0]), "idle_between_issues": med(iss[:, ev[1:], 0] - iss[:, ev[1:] - 1, 1])}
lat_a, lat_acc = [], []
for s in range(slots):
    for e in hid:
        nxt = (e + 1) * slots + s
        if nxt < iss.shape[1]:
            lat_a.append(iss[:, nxt, 0] - ep[:, s, e, 4])
        cur = e * slots + s
        if cur < iss.shape[1]:
            lat_acc.append(ep[:, s, e, 1] - iss[:, cur, 1])
res["handoff"] = {"a_ready_to_issue_start": med(np.concatenate(lat_a)) if lat_a else None, "commit_to_accumulator_seen (MMA drain + barrier)": med(np.concatenate(lat_acc)) if lat_acc else None}
res["mma_floor_cycles_per_layer"] = (width * width) / 32
# raw timeline of one CTA (cycles relative to its first stamp): issuer events [slot][layer-event][fields], epilogue events per slot
cta = c[3]
t0 = np.nanmin(np.where(cta > 0, cta, np.nan))
raw = {"issuer": [[int(v - t0) if v > 0 else None for v in cta[0, e, :2]] for e in range(min(40, iss.shape[1]))]}
for s_ in range(slots):
    raw[f"slot{s_}"] = [[int(v - t0) if v > 0 else None for v in cta[1 + s_, e, :5]] for e in range(min(20, n_ev))]
res["raw_cta3"] = raw
print(json.dumps(res))
