"""Diagnostic (not a test): run one fused training step with all taps and print, stage by stage, how the device compares
with the oracle. Usage on the GPU box: python tests/debug_stages.py [config-name] [n_in] [n_out] [B]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-cuda-nn_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import oracle_binding as ob
import tcnn_b200


def stats(name, dev, ref):
    dev = np.asarray(dev, np.float64)
    ref = np.asarray(ref, np.float64)
    d = np.abs(dev - ref)
    denom = np.abs(ref).max() + 1e-30
    print(f"  {name:18s} shape={dev.shape} max|ref|={np.abs(ref).max():.4e} max|dev|={np.abs(dev).max():.4e} max|d|={d.max():.3e} "
          f"rel={d.max() / denom:.3e} mismatch={(dev != ref).mean():.4f} nan={np.isnan(dev).sum()}")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "hash3d_small"
    n_in = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n_out = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    B = int(sys.argv[4]) if len(sys.argv) > 4 else 512
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "configs", name + ".json")))
    NH = cfg["network"]["n_hidden_layers"]
    model = tcnn_b200.create_from_config(n_in, n_out, cfg)
    lv = model.grid_levels()
    orc = ob.OracleModel(n_in, n_out, cfg, scales=lv["scales"])
    print("levels: dev scales", [float(np.float32(s)) for s in lv["scales"]])
    print("        host scales", [float(orc.grid.scales[i]) for i in range(orc.grid.n_levels)] if False else "(oracle uses device scales)")
    host = ob.OracleModel(n_in, n_out, cfg)
    print("        host-vs-dev scale bits equal:", [bool(np.float32(host.grid.scales[i]) == np.float32(lv["scales"][i])) for i in range(host.grid.n_levels)])
    p0 = model.trainer.params_full_precision().cpu().numpy()
    print("init params bit-equal:", np.array_equal(p0.view(np.uint32), orc.params_fp32.view(np.uint32)),
          "mlp", np.array_equal(p0[:orc.n_mlp], orc.params_fp32[:orc.n_mlp]), "max|d|", np.abs(p0 - orc.params_fp32).max())
    rng = ob.default_rng(1337)
    x = ob.generate_random_uniform(rng, B * n_in).reshape(B, n_in)
    y = ob.make_targets(x, n_out)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    taps = dict(
        encoded=torch.zeros(B, 64, dtype=torch.float16, device="cuda"), hidden=torch.zeros(NH, B, 64, dtype=torch.float16, device="cuda"),
        output=torch.zeros(B, 16, dtype=torch.float16, device="cuda"), dL_doutput=torch.zeros(B, 16, dtype=torch.float16, device="cuda"),
        grad_hidden=torch.zeros(NH, B, 64, dtype=torch.float16, device="cuda"), dL_dencoded=torch.zeros(B, 64, dtype=torch.float16, device="cuda"),
        loss_values=torch.zeros(B, n_out, dtype=torch.float32, device="cuda"))
    model.set_debug_taps(**taps)
    model.trainer.training_step(xd, yd, run_optimizer=False)
    loss = model.trainer.loss()
    torch.cuda.synchronize()
    print("step done, loss", loss)
    f16 = lambda t: t.cpu().numpy().view(np.uint16)
    h2f = ob.half_bits_to_float
    W = orc.grid.padded_width
    enc_dev = f16(taps["encoded"])
    enc_ref = orc.encode(x)
    stats("encoded", h2f(enc_dev[:, :W].T), h2f(enc_ref))
    bad = np.argwhere(enc_dev[:, :W].T != enc_ref)
    if len(bad):
        print("   first mismatches (feature,sample):", bad[:8].tolist(), "per-level mismatch:", [(int((enc_dev[:, 2*l:2*l+2].T != enc_ref[2*l:2*l+2]).sum())) for l in range(orc.grid.n_levels)])
    hid_ref, out_ref = orc.mlp_forward(np.ascontiguousarray(enc_dev[:, :W].T))
    hid_dev = f16(taps["hidden"])
    for l in range(NH):
        stats(f"hidden[{l}]", h2f(hid_dev[l]), h2f(hid_ref[l]))
    # also: hidden computed from device hidden[l-1] to isolate layers
    out_dev = f16(taps["output"])
    stats("output", h2f(out_dev), h2f(out_ref))
    print("   out_dev[0,:4]", h2f(out_dev[0, :4]), "out_ref[0,:4]", h2f(out_ref[0, :4]))
    lv_ref, dy_ref = orc.loss(out_dev, y)
    stats("loss_values", taps["loss_values"].cpu().numpy(), lv_ref[:, :n_out])
    dy_dev = f16(taps["dL_doutput"])
    stats("dL_doutput", h2f(dy_dev), h2f(dy_ref))
    dW_ref, denc_ref = orc.mlp_backward(np.ascontiguousarray(enc_dev[:, :W].T), hid_dev, dy_dev)
    # hidden gradients: recompute reference chain pieces through oracle is internal; report device magnitudes
    gh = h2f(f16(taps["grad_hidden"]))
    print("   grad_hidden max|.| per layer:", [float(np.abs(gh[l]).max()) for l in range(NH)], "nan:", int(np.isnan(gh).sum()))
    denc_dev = f16(taps["dL_dencoded"])
    stats("dL_dencoded", h2f(denc_dev[:, :W].T), h2f(denc_ref))
    grads = h2f(f16(model.trainer.param_gradients()))
    n_mlp = orc.n_mlp
    in_w, NHm = W, NH
    offs = [0, 64 * in_w] + [64 * in_w + (i + 1) * 4096 for i in range(NH - 1)]
    offs.append(n_mlp)
    for i in range(len(offs) - 1):
        stats(f"dW[{i}]", grads[offs[i]:offs[i + 1]], dW_ref[offs[i]:offs[i + 1]].astype(np.float16).astype(np.float64))
    g_ref = orc.grid_backward(x, np.ascontiguousarray(denc_dev[:, :W].T))
    stats("grid grads", grads[n_mlp:], g_ref.astype(np.float16).astype(np.float64))
    for l in range(orc.grid.n_levels):
        a, b = orc.grid.offsets[l] * 2, orc.grid.offsets[l + 1] * 2
        d = np.abs(grads[n_mlp + a:n_mlp + b] - g_ref[a:b])
        print(f"     level {l:2d}: max|ref|={np.abs(g_ref[a:b]).max():.3e} max|d|={d.max():.3e} nonzero dev/ref = {(grads[n_mlp+a:n_mlp+b]!=0).sum()}/{(g_ref[a:b]!=0).sum()}")
    model.trainer.training_step(xd, yd)
    l2 = model.trainer.loss()
    lr = orc.training_step(x, y)
    p1 = model.trainer.params_full_precision().cpu().numpy()
    stats("params after step", p1, orc.params_fp32)
    print("loss dev", l2, "oracle", lr)
    out = model.network.inference(xd).cpu().numpy()
    stats("inference", out, orc.inference(x))


if __name__ == "__main__":
    main()
