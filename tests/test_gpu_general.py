"""GPU parity of the GENERAL (unfused) path: configurations the fused kernel does not cover -- 128 neurons (BASELINE.json configs[2]),
n_features_per_level in {1, 4, 8}, 4-D inputs, Nearest interpolation, more than 16 outputs -- run as encoding kernel -> stand-alone MLP
kernels (forward, dgrad chain with the weights read transposed in place, weight-gradient kernel) -> encoding backward kernel.

Checked against the CPU oracle stage by stage (network backward: orc_mlp_backward, fully_fused_mlp.cu:733-866) and as whole training
trajectories (orc_training_step, trainer.h:254-357), with the bars of tests/test_gpu_parity.py."""
import ctypes
import json

import numpy as np
import pytest

import oracle_binding as ob
from golden_util import rae

pytestmark = pytest.mark.gpu


def f16(t):
    return t.cpu().numpy().view(np.uint16)


def oracle_mlp(width, n_hidden, in_w, out_pad, act, out_act, weights16):
    lib = ob.load()
    mlp = ob.Mlp(in_w, width, n_hidden, out_pad, 0, ob.ACT[act.lower()], ob.ACT[out_act.lower()], 0)
    assert lib.orc_mlp_setup(ctypes.byref(mlp)) == 0
    assert mlp.n_params == weights16.size
    return lib, mlp


BACKWARD_SHAPES = [
    # width, hidden, n_in, n_out, activation, batch
    (128, 4, 128, 128, "ReLU", 1024),
    (128, 4, 32, 16, "ReLU", 148 * 128 * 3 + 256),   # several tiles per CTA, contiguous tile ranges of different lengths
    (128, 2, 64, 48, "Tanh", 1024),
    (128, 8, 128, 16, "ReLU", 2048),                 # 9 matrices: ring mode in the dgrad chain, three launches of the weight-gradient kernel
    (64, 2, 64, 64, "ReLU", 148 * 128 * 2 + 512),
    (64, 4, 32, 16, "LeakyReLU", 2048),
    (64, 9, 64, 16, "ReLU", 1024),                   # 10 matrices > 8 accumulators: two weight-gradient launches
    (32, 3, 32, 32, "ReLU", 2048),
    (16, 2, 16, 16, "Sigmoid", 1024),
    (16, 3, 32, 16, "ReLU", 1024),                   # input wider than the layers
]


@pytest.mark.parametrize("width,n_hidden,n_in,n_out,act,B", BACKWARD_SHAPES)
def test_network_backward_matches_oracle(torch_cuda, width, n_hidden, n_in, n_out, act, B):
    torch = torch_cuda
    import tcnn_b200

    cfg = {"otype": "FullyFusedMLP", "n_neurons": width, "n_hidden_layers": n_hidden, "activation": act, "output_activation": "None"}
    net = tcnn_b200.Network(n_in, n_out, cfg)
    out_pad = net.padded_output_width
    p16 = net.initial_params(seed=3).to(torch.float16).contiguous()
    g = torch.Generator(device="cuda").manual_seed(7)
    x = (torch.rand(B, n_in, device="cuda", generator=g) * 2 - 1).to(torch.float16).contiguous()
    dy = ((torch.rand(B, out_pad, device="cuda", generator=g) * 2 - 1) * 0.25).to(torch.float16).contiguous()
    out, hidden = net.forward(x, p16)
    dx, dp = net.backward(x, out, hidden, dy, p16)
    dx_only, none = net.backward(x, out, hidden, dy, p16, want_param_grad=False)
    torch.cuda.synchronize()
    assert none is None and torch.equal(dx_only, dx)

    lib, mlp = oracle_mlp(width, n_hidden, n_in, out_pad, act, "None", f16(p16))
    # dL/d(input): per-sample, so a head + tail sample of the batch is enough for the scalar oracle
    n_check = min(B, 2048)
    sel = np.r_[0 : n_check // 2, B - n_check // 2 : B]
    x_soa = np.ascontiguousarray(f16(x)[sel].T)
    hid_dev = np.ascontiguousarray(f16(hidden)[:, sel])
    dW_sel = np.zeros(net.n_params, np.float64)
    dx_ref = np.zeros((n_in, len(sel)), np.uint16)
    lib.orc_mlp_backward(ctypes.byref(mlp), len(sel), ob.ACCUM_FP32, ob._p(f16(p16)), ob._p(x_soa), ob._p(hid_dev), ob._p(np.ascontiguousarray(f16(dy)[sel])), ob._p(dW_sel), ob._p(dx_ref))
    a, b = ob.half_bits_to_float(f16(dx)[sel]), ob.half_bits_to_float(dx_ref.T)
    assert rae(a, b, 99.9) < (2e-3 if act in ("ReLU", "LeakyReLU") else 5e-3), rae(a, b, 99.9)
    # weight gradients: sums over the WHOLE batch -> oracle over the whole batch when it is small, else check linearity below
    if B <= 4096:
        dW = np.zeros(net.n_params, np.float64)
        lib.orc_mlp_backward(ctypes.byref(mlp), B, ob.ACCUM_FP32, ob._p(f16(p16)), ob._p(np.ascontiguousarray(f16(x).T)), ob._p(np.ascontiguousarray(f16(hidden))), ob._p(f16(dy)), ob._p(dW), None)
        got = dp.float().cpu().numpy()
        want = dW.astype(np.float16).astype(np.float32)
        assert rae(got, want, 99.9) < 3e-3, rae(got, want, 99.9)
        assert np.abs(got - want).max() <= 2e-2 * np.abs(want).max()
    else:
        # split the batch in two multiples of 256: gradients add up (every tile range / launch split is exercised by the big call)
        h = (B // 2) // 256 * 256
        parts = []
        for lo, hi in ((0, h), (h, B)):
            o2, hid2 = net.forward(x[lo:hi].contiguous(), p16)
            parts.append(net.backward(x[lo:hi].contiguous(), o2, hid2, dy[lo:hi].contiguous(), p16, want_input_grad=False)[1].float())
        total = (parts[0] + parts[1]).cpu().numpy()
        got = dp.float().cpu().numpy()
        assert rae(got, total, 99.9) < 2e-3


def test_network_backward_through_output_activation_and_module_tier(torch_cuda):
    """cpp::create_network's Module::backward: fp32 inputs through the Identity encoding, dL/d(input) in fp32, an output activation."""
    torch = torch_cuda
    import tcnn_b200

    cfg = {"otype": "FullyFusedMLP", "n_neurons": 64, "n_hidden_layers": 2, "activation": "ReLU", "output_activation": "Sigmoid"}
    net = tcnn_b200.Network(3, 4, cfg)
    B = 1024
    p16 = net.initial_params(seed=5).to(torch.float16).contiguous()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(B, 3, device="cuda", generator=g)
    dy = ((torch.rand(B, 16, device="cuda", generator=g) * 2 - 1)).to(torch.float16).contiguous()
    dx, dp = net.module_backward(x, dy, p16)
    torch.cuda.synchronize()
    # reference chain in torch fp32 on the same fp16 weights (Identity encoding: padding features are ONE)
    w = p16.float()
    W0, W1, Wo = w[: 64 * 16].view(64, 16), w[64 * 16 : 64 * 16 + 64 * 64].view(64, 64), w[64 * 16 + 64 * 64 :].view(16, 64)
    xin = torch.cat([x.half().float(), torch.ones(B, 13, device="cuda")], 1).requires_grad_(True)
    h0 = torch.relu(xin @ W0.T).half().float()
    W0g, W1g, Wog = W0.clone().requires_grad_(True), W1.clone().requires_grad_(True), Wo.clone().requires_grad_(True)
    h0 = torch.relu(xin @ W0g.T)
    h1 = torch.relu(h0 @ W1g.T)
    y = torch.sigmoid(h1 @ Wog.T)
    y.backward(dy.float())
    want_dp = torch.cat([W0g.grad.flatten(), W1g.grad.flatten(), Wog.grad.flatten()]).cpu().numpy()
    assert rae(dp.float().cpu().numpy(), want_dp, 99.0) < 2e-2
    assert rae(dx.cpu().numpy(), xin.grad[:, :3].cpu().numpy(), 99.0) < 2e-2
    out = net.module_inference(x, p16)
    assert rae(out.float().cpu().numpy()[:, :4], y.detach().cpu().numpy()[:, :4], 99.0) < 5e-3


GENERAL_CONFIGS = {
    # BASELINE.json configs[2]: samples/mlp_learning_an_image.cu's model -- 2-D hash grid + 128 x 4 FullyFusedMLP
    "image_w128": (2, 3, 512, {"encoding": {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 15, "base_resolution": 16, "per_level_scale": 1.5},
                               "network": {"otype": "FullyFusedMLP", "n_neurons": 128, "n_hidden_layers": 4, "activation": "ReLU", "output_activation": "None"}}),
    "f4_3d": (3, 3, 512, {"encoding": {"otype": "HashGrid", "n_levels": 8, "n_features_per_level": 4, "log2_hashmap_size": 14, "base_resolution": 8, "per_level_scale": 1.5},
                          "network": {"otype": "FullyFusedMLP", "n_neurons": 64, "n_hidden_layers": 2}}),
    "f8_2d": (2, 1, 256, {"encoding": {"otype": "HashGrid", "n_levels": 4, "n_features_per_level": 8, "log2_hashmap_size": 12, "base_resolution": 8, "per_level_scale": 2.0},
                          "network": {"otype": "FullyFusedMLP", "n_neurons": 32, "n_hidden_layers": 2}}),
    "f1_3d": (3, 2, 256, {"encoding": {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 1, "log2_hashmap_size": 13, "base_resolution": 4, "per_level_scale": 1.5},
                          "network": {"otype": "FullyFusedMLP", "n_neurons": 16, "n_hidden_layers": 3}}),
    "d4": (4, 3, 256, {"encoding": {"otype": "HashGrid", "n_levels": 8, "n_features_per_level": 2, "log2_hashmap_size": 14, "base_resolution": 4, "per_level_scale": 1.5},
                       "network": {"otype": "FullyFusedMLP", "n_neurons": 64, "n_hidden_layers": 2}}),
    "nearest": (3, 3, 256, {"encoding": {"otype": "HashGrid", "n_levels": 8, "n_features_per_level": 2, "log2_hashmap_size": 14, "base_resolution": 8, "per_level_scale": 1.5,
                                         "interpolation": "Nearest"},
                            "network": {"otype": "FullyFusedMLP", "n_neurons": 64, "n_hidden_layers": 2}}),
    "wide_out": (3, 20, 256, {"encoding": {"otype": "HashGrid", "n_levels": 8, "n_features_per_level": 2, "log2_hashmap_size": 14, "base_resolution": 8, "per_level_scale": 1.5},
                              "network": {"otype": "FullyFusedMLP", "n_neurons": 64, "n_hidden_layers": 2}}),
}
for _c in GENERAL_CONFIGS.values():
    _c[3].setdefault("loss", {"otype": "RelativeL2"})
    _c[3].setdefault("optimizer", {"otype": "Adam", "learning_rate": 1e-2, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6})


def make_batch(n_in, n_out, B, seed=1337):
    rng = ob.default_rng(seed)
    x = ob.generate_random_uniform(rng, B * n_in).reshape(B, n_in)
    return x, ob.make_targets(x, n_out)


@pytest.mark.parametrize("name", list(GENERAL_CONFIGS))
def test_general_path_step_matches_oracle(torch_cuda, name):
    torch = torch_cuda
    import tcnn_b200

    n_in, n_out, B, cfg = GENERAL_CONFIGS[name]
    model = tcnn_b200.create_from_config(n_in, n_out, cfg)
    orc = ob.OracleModel(n_in, n_out, cfg, scales=model.grid_levels()["scales"])
    assert model.n_params == orc.n_params
    p0 = model.trainer.params_full_precision().cpu().numpy()
    assert np.array_equal(p0, orc.params_fp32)  # same initialisation stream
    x, y = make_batch(n_in, n_out, B)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()

    # one step without the optimizer: loss and every gradient
    model.trainer.training_step(xd, yd, run_optimizer=False)
    loss_dev = model.trainer.loss()
    grads = model.trainer.param_gradients().float().cpu().numpy()
    loss_ref = orc.training_step(x, y, run_optimizer=False)
    gref = ob.half_bits_to_float(orc.grads_fp16)
    assert abs(loss_dev - loss_ref) <= 1e-3 * abs(loss_ref)
    n_mlp = orc.n_mlp
    assert rae(grads[:n_mlp], gref[:n_mlp], 99.9) < 1.2e-2
    touched_dev, touched_ref = grads[n_mlp:] != 0, gref[n_mlp:] != 0
    assert (touched_dev != touched_ref).mean() < 2e-3
    assert rae(grads[n_mlp:], gref[n_mlp:], 99.9) < 1.2e-2
    out_dev = model.network.inference(xd).cpu().numpy()
    assert rae(out_dev, orc.inference(x), 99.0) < 1e-2

    # ten optimiser steps: the trajectory
    dev_losses, ref_losses = [], []
    for _ in range(10):
        model.trainer.training_step(xd, yd)
        dev_losses.append(model.trainer.loss())
        ref_losses.append(orc.training_step(x, y))
    assert dev_losses[-1] < dev_losses[0]
    for a, b in zip(dev_losses, ref_losses):
        assert abs(a - b) <= 3e-2 * abs(b) + 1e-6, (dev_losses, ref_losses)


def test_general_path_module_tier(torch_cuda):
    """tcnn::cpp::Module of a 128-neuron NetworkWithInputEncoding: forward == inference, backward == the trainer tier's gradients, and
    dL/d(input) against the oracle."""
    torch = torch_cuda
    import tcnn_b200

    n_in, n_out, B, cfg = GENERAL_CONFIGS["image_w128"]
    mod = tcnn_b200.Module(n_in, n_out, cfg["encoding"], cfg["network"])
    p32 = mod.initial_params(seed=1337)
    p16 = p32.to(torch.float16).contiguous()
    x, _ = make_batch(n_in, n_out, B)
    xd = torch.from_numpy(x).cuda()
    g = torch.Generator(device="cuda").manual_seed(2)
    dy = ((torch.rand(B, 16, device="cuda", generator=g) * 2 - 1) * 0.1).to(torch.float16).contiguous()
    out = mod.fwd(xd, p16)
    dp, dx = mod.bwd(xd, p16, dy, want_input_grad=True)
    torch.cuda.synchronize()
    scales = tcnn_b200.create_from_config(n_in, n_out, cfg).grid_levels()["scales"]  # device-evaluated level scales
    orc = ob.OracleModel(n_in, n_out, cfg, scales=scales)
    orc.set_params_full_precision(p16.float().cpu().numpy())
    want = orc.backward_from_dy(x, f16(dy))
    got = dp.float().cpu().numpy()
    assert rae(got[: orc.n_mlp], want[: orc.n_mlp], 99.9) < 1.2e-2
    assert rae(got[orc.n_mlp :], want[orc.n_mlp :], 99.9) < 1.2e-2
    enc = orc.encode(x)
    hidden, out_ref = orc.mlp_forward(enc)
    assert rae(out.float().cpu().numpy(), ob.half_bits_to_float(out_ref), 99.0) < 1e-2
    _, d_enc = orc.mlp_backward(enc, hidden, np.ascontiguousarray(f16(dy)))
    dx_ref = orc.grid_input_gradient(x, d_enc)
    assert rae(dx.cpu().numpy(), dx_ref, 99.0) < 2e-2


def test_general_path_rejects_what_no_kernel_covers(torch_cuda):
    import tcnn_b200

    cfg = json.loads(json.dumps(GENERAL_CONFIGS["f4_3d"][3]))
    cfg["encoding"]["n_features_per_level"] = 3
    with pytest.raises(tcnn_b200.TcnnError, match="n_features_per_level must be 1, 2, 4, or 8"):
        tcnn_b200.create_from_config(3, 3, cfg)
    cfg = json.loads(json.dumps(GENERAL_CONFIGS["f4_3d"][3]))
    cfg["encoding"]["n_levels"] = 16  # 64 features into 32 neurons: dL/d(encoded) is wider than the stand-alone chain holds
    cfg["network"]["n_neurons"] = 32
    with pytest.raises(tcnn_b200.TcnnError):
        tcnn_b200.create_from_config(3, 3, cfg)


@pytest.mark.parametrize("name", ["hash3d_small", "image2d", "identity_cutlass"])
def test_fused_and_general_paths_agree(torch_cuda, name):
    """The same configuration through the fused kernel and (tcnnb_debug_set("general", 1)) through the stand-alone kernels: identical
    encodings and fp32-accumulated layers, so outputs agree to fp16 rounding and gradients to the order of the atomics."""
    torch = torch_cuda
    import tcnn_b200
    from golden_util import GOLDEN

    cfg = json.load(open(f"{GOLDEN}/configs/{name}.json"))
    n_in = 2 if name == "image2d" else 3
    B = 2048
    model = tcnn_b200.create_from_config(n_in, 3, cfg)
    x, y = make_batch(n_in, 3, B)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    res = {}
    for path in ("fused", "general"):
        model.debug_set("general", 1 if path == "general" else 0)
        out = model.network.inference(xd).clone()
        model.trainer.training_step(xd, yd, run_optimizer=False)
        res[path] = (out, model.trainer.loss(), model.trainer.param_gradients().float().clone())
    (o1, l1, g1), (o2, l2, g2) = res["fused"], res["general"]
    assert float((o1 - o2).abs().max()) <= 2e-3 * max(1.0, float(o1.abs().max()))
    assert abs(l1 - l2) <= 1e-3 * abs(l1)
    n_mlp = model.n_mlp_params
    assert rae(g1[:n_mlp].cpu().numpy(), g2[:n_mlp].cpu().numpy(), 99.9) < 5e-3
    if model.n_params > n_mlp:
        a, b = g1[n_mlp:].cpu().numpy(), g2[n_mlp:].cpu().numpy()
        assert ((a != 0) != (b != 0)).mean() < 1e-3
        assert rae(a, b, 99.9) < 1e-2
    # and a few optimiser steps on the general path keep training the same model
    losses = []
    for _ in range(5):
        model.trainer.training_step(xd, yd)
        losses.append(model.trainer.loss())
    assert losses[-1] < losses[0]


def test_general_path_full_size_properties(torch_cuda):
    """BASELINE.json configs[2] (2-D HashGrid + 128 x 4, the model of samples/mlp_learning_an_image.cu) at the benchmarked batch 2^18 on the
    general path -- 13.8 tiles per CTA in the network kernels, replicated scatter of the coarse grid levels, one packed weight-gradient
    launch: batch linearity (gradients of the batch == sum over its two halves normalised over the whole batch), zero-gradient entries
    skipped by Adam, the loss decreases, host-buffer step == device step."""
    torch = torch_cuda
    import tcnn_b200

    n_in, n_out, _, cfg = GENERAL_CONFIGS["image_w128"]
    B = 1 << 18
    model = tcnn_b200.create_from_config(n_in, n_out, cfg)
    x, y = make_batch(n_in, n_out, B)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    model.trainer.training_step(xd, yd, run_optimizer=False)
    full_loss = model.trainer.loss()
    g_full = model.trainer.param_gradients().float().clone()
    h = B // 2
    model.trainer.training_step_shard(xd[:h], yd[:h], B)
    la, ga = model.trainer.loss(), model.trainer.param_gradients().float().clone()
    model.trainer.training_step_shard(xd[h:], yd[h:], B)
    lb, gb = model.trainer.loss(), model.trainer.param_gradients().float().clone()
    assert abs((la + lb) - full_loss) <= 1e-3 * full_loss
    gsum = ga + gb
    nz = g_full != 0
    assert ((gsum != 0) == nz).float().mean() > 0.999
    assert rae(gsum[nz].cpu().numpy(), g_full[nz].cpu().numpy(), 99.0) < 2e-2
    n_mlp = model.n_mlp_params
    p_before = model.trainer.params_full_precision().clone()
    losses = []
    for _ in range(6):
        model.trainer.training_step(xd, yd)
        losses.append(model.trainer.loss())
    p_after = model.trainer.params_full_precision()
    assert losses[-1] < 0.5 * losses[0] and torch.isfinite(p_after).all()
    assert (p_before[:n_mlp] != p_after[:n_mlp]).float().mean() > 0.99
    l_host = model.training_step_host(x, y)
    assert np.isfinite(l_host) and l_host < losses[0]
    out = model.network.inference(xd)
    assert torch.isfinite(out).all() and out.shape == (B, n_out)
