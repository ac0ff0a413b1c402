"""general_split.py -- where the general (unfused) training step spends its time: the stages timed one by one through the library's own
tiers (tcnn_b200.Encoding / tcnn_b200.Network), CUDA events, device-resident data.

    python scripts/general_split.py [image_w128|f4_l8|...] [iters]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-cuda-nn_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def timed(torch, fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    import torch

    import tcnn_b200
    from bench_configs import CONFIGS

    name = sys.argv[1] if len(sys.argv) > 1 else "image_w128"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    n_in, n_out, B, cfg, what = CONFIGS[name]
    enc = tcnn_b200.Encoding(n_in, cfg["encoding"])
    width, hidden = cfg["network"]["n_neurons"], cfg["network"]["n_hidden_layers"]
    in_w = (enc.n_output_dims + 15) // 16 * 16
    net = tcnn_b200.Network(in_w, n_out, cfg["network"])
    x = torch.rand(B, n_in, device="cuda")
    table16 = enc.initial_params(1337).to(torch.float16)
    p16 = net.initial_params(7).to(torch.float16)
    feats = enc.fwd(x, table16)
    if feats.shape[1] != in_w:
        feats = torch.nn.functional.pad(feats, (0, in_w - feats.shape[1]))
    feats = feats.contiguous()
    out, hid = net.forward(feats, p16)
    dy = (torch.rand(B, net.padded_output_width, device="cuda") * 0.01).to(torch.float16)
    denc, _ = net.backward(feats, out, hid, dy, p16, want_param_grad=False)
    res = {"config": name, "batch": B, "width": width, "n_hidden_layers": hidden, "encoded_width": in_w}
    res["encoding_forward_ms"] = timed(torch, lambda: enc.fwd(x, table16), iters)
    res["mlp_inference_ms"] = timed(torch, lambda: net.inference_mixed_precision(feats, p16), iters)
    res["mlp_forward_keep_hidden_ms"] = timed(torch, lambda: net.forward(feats, p16), iters)
    res["mlp_dgrad_chain_ms"] = timed(torch, lambda: net.backward(feats, out, hid, dy, p16, want_param_grad=False), iters)
    res["mlp_dgrad_no_dinput_plus_wgrad_ms"] = timed(torch, lambda: net.backward(feats, out, hid, dy, p16, want_input_grad=False), iters)
    res["mlp_backward_all_ms"] = timed(torch, lambda: net.backward(feats, out, hid, dy, p16), iters)
    dfe = denc[:, : enc.n_output_dims].contiguous() if denc.shape[1] != enc.n_output_dims else denc
    res["encoding_backward_ms"] = timed(torch, lambda: enc.bwd(x, table16, dfe), iters)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
