"""GPU parity of the stand-alone grid-encoding kernels (csrc/grid_kernels.cu, C ABI tcnnb_encoding_*) and of the module tier's
input-position gradient against the CPU oracle: the reference's whole grid configuration space -- n_features_per_level 1/2/4/8,
2-4 input dimensions, Hash/Dense/Tiled, Nearest/Linear/Smoothstep (grid.h:1757-1768,1824-1834). Encoded features are bit-exact
(integer indices + the reference's fp16 fma chain); gradients within the reference's bars (tests/test_common.h:216-218)."""
import json
import os

import numpy as np
import pytest

import oracle_binding as ob
from golden_util import rae

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def f16(t):
    return t.cpu().numpy().view(np.uint16)


def make_x(n_in, B, seed=31):
    rng = ob.default_rng(seed)
    return ob.generate_random_uniform(rng, B * n_in).reshape(B, n_in)


CONFIGS = [
    # (n_input_dims, encoding config)
    (3, {"otype": "HashGrid", "n_levels": 8, "n_features_per_level": 1, "log2_hashmap_size": 12, "base_resolution": 8, "per_level_scale": 1.5}),
    (3, {"otype": "HashGrid", "n_levels": 8, "n_features_per_level": 2, "log2_hashmap_size": 12, "base_resolution": 8, "per_level_scale": 1.5}),
    (3, {"otype": "HashGrid", "n_levels": 8, "n_features_per_level": 4, "log2_hashmap_size": 12, "base_resolution": 8, "per_level_scale": 1.5}),
    (3, {"otype": "HashGrid", "n_levels": 6, "n_features_per_level": 8, "log2_hashmap_size": 11, "base_resolution": 4, "per_level_scale": 2.0}),
    (2, {"otype": "HashGrid", "n_levels": 10, "n_features_per_level": 4, "log2_hashmap_size": 10, "base_resolution": 16, "per_level_scale": 1.5}),
    (4, {"otype": "HashGrid", "n_levels": 6, "n_features_per_level": 2, "log2_hashmap_size": 12, "base_resolution": 4, "per_level_scale": 1.5}),
    (4, {"otype": "HashGrid", "n_levels": 4, "n_features_per_level": 1, "log2_hashmap_size": 10, "base_resolution": 3, "per_level_scale": 2.0}),
    (3, {"otype": "DenseGrid", "n_levels": 4, "n_features_per_level": 4, "base_resolution": 4, "per_level_scale": 1.5}),
    (3, {"otype": "TiledGrid", "n_levels": 5, "n_features_per_level": 2, "base_resolution": 5, "per_level_scale": 1.5}),
    (3, {"otype": "HashGrid", "n_levels": 8, "n_features_per_level": 2, "log2_hashmap_size": 12, "base_resolution": 8, "per_level_scale": 1.5, "interpolation": "Nearest"}),
    (3, {"otype": "HashGrid", "n_levels": 8, "n_features_per_level": 4, "log2_hashmap_size": 12, "base_resolution": 8, "per_level_scale": 1.5, "interpolation": "Smoothstep"}),
    (2, {"otype": "HashGrid", "n_levels": 8, "n_features_per_level": 1, "log2_hashmap_size": 12, "base_resolution": 8, "per_level_scale": 1.5, "interpolation": "Nearest"}),
]


@pytest.mark.parametrize("n_in,enc_cfg", CONFIGS, ids=[f"D{n}_F{c['n_features_per_level']}_{c['otype']}_{c.get('interpolation', 'Linear')}" for n, c in CONFIGS])
def test_encoding_matches_oracle(torch_cuda, n_in, enc_cfg):
    torch = torch_cuda
    import tcnn_b200

    enc = tcnn_b200.Encoding(n_in, enc_cfg)
    levels = enc.grid_levels()
    orc = ob.OracleModel(n_in, 1, {"encoding": enc_cfg, "network": {"n_neurons": 16, "n_hidden_layers": 1}}, scales=levels["scales"])
    F, L = enc_cfg["n_features_per_level"], enc_cfg["n_levels"]
    assert enc.n_output_dims == L * F and enc.n_params == orc.grid.n_params
    assert levels["offsets"] == list(orc.grid.offsets[: L + 1])
    p32 = enc.initial_params(seed=9, scale=50.0)  # larger than the default 1e-4 range so that fp16 features are well above denormals
    assert float(p32.abs().max()) <= 5e-3 + 1e-9 and float(p32.abs().max()) > 2.5e-3
    p16 = p32.to(torch.float16).contiguous()
    B = 2048
    x = make_x(n_in, B)
    xd = torch.from_numpy(x).cuda()

    # ---- forward: bit-exact
    out = enc.fwd(xd, p16)
    torch.cuda.synchronize()
    orc.params_fp16[orc.n_mlp :] = f16(p16)
    enc_ref = orc.encode(x)  # SoA [padded][B]
    assert np.array_equal(f16(out).T, enc_ref[: L * F]), "encoded features differ"

    # ---- backward: parameter gradients (fp16 atomics vs exact sums) and input gradients (fp32, same summation order)
    g = torch.Generator(device="cuda").manual_seed(1)
    dy = (torch.randn(B, L * F, device="cuda", generator=g) * 0.05).to(torch.float16).contiguous()
    gp, gx = enc.bwd(xd, p16, dy, want_params=True, want_input=True)
    torch.cuda.synchronize()
    dy_soa = np.zeros((orc.grid.padded_width, B), np.uint16)
    dy_soa[: L * F] = f16(dy).T
    g_ref = orc.grid_backward(x, dy_soa)
    g_dev = ob.half_bits_to_float(f16(gp)).astype(np.float64)
    assert rae(g_dev, g_ref.astype(np.float16).astype(np.float64), 99.9) < 1.2e-2
    assert (g_dev[g_ref == 0] == 0).all() or ((g_dev != 0) & (g_ref == 0)).mean() < 1e-4
    gx_ref = orc.grid_input_gradient(x, dy_soa)
    gx_dev = gx.cpu().numpy()
    if enc_cfg.get("interpolation") == "Nearest":
        assert (gx_dev == 0).all() and (gx_ref == 0).all()
    else:
        assert rae(gx_dev, gx_ref, 99.0) < 1e-3  # tests/test_common.h:216 asks 1e-2 of input gradients
        assert np.abs(gx_dev - gx_ref).max() <= 1e-3 * np.abs(gx_ref).max()
    # gradients are overwritten, not accumulated
    gp2, _ = enc.bwd(xd, p16, dy, want_params=True, want_input=False)
    assert rae(ob.half_bits_to_float(f16(gp2)), g_dev, 99.9) < 5e-3

    # ---- max_level (GridEncoding::set_max_level, grid.h:69-92): masked levels encode to zero, the others are unchanged
    enc.set_max_level(0.5)
    half = enc.fwd(xd, p16)
    torch.cuda.synchronize()
    n_active = min(L, int(np.floor(0.5 * L + 1e-3)) + 1)  # level l is active while l < max_level * L + 1e-3 (grid.h:75)
    assert np.array_equal(f16(half)[:, : n_active * F], f16(out)[:, : n_active * F])
    assert (f16(half)[:, n_active * F :] & 0x7FFF == 0).all()


def test_module_input_gradient_matches_oracle(torch_cuda):
    """tcnn::cpp::Module::backward with dL_dinput (src/cpp_api.cu:104-125) on the fused module tier: the network's input gradient
    rows contracted with d(encoded)/d(position) (grid.h:170-212,322-350), against the oracle's stage functions."""
    torch = torch_cuda
    import tcnn_b200

    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "configs", "hash3d_small.json")))
    for interp in ("Linear", "Smoothstep"):
        cfg["encoding"]["interpolation"] = interp
        mod = tcnn_b200.Module(3, 3, cfg["encoding"], cfg["network"])
        model = tcnn_b200.create_from_config(3, 3, cfg)
        orc = ob.OracleModel(3, 3, cfg, scales=model.grid_levels()["scales"])
        # parameters large enough that the encoding matters
        p32 = mod.initial_params(seed=4)
        p32[model.n_mlp_params :] *= 200.0
        p16 = p32.to(torch.float16).contiguous()
        orc.params_fp16[:] = f16(p16)
        B = 1024
        x = make_x(3, B, seed=8)
        xd = torch.from_numpy(x).cuda()
        rng = np.random.default_rng(5)
        dy = (rng.standard_normal((B, 16)) * 0.5).astype(np.float16)
        dy[:, 3:] = 0
        dyd = torch.from_numpy(dy).cuda()
        gp, gx = mod.bwd(xd, p16, dyd, want_input_grad=True)
        _, gx_only = mod.bwd(xd, p16, dyd, want_input_grad=True, want_param_grad=False)
        torch.cuda.synchronize()
        assert torch.equal(gx, gx_only)
        enc = orc.encode(x)
        hidden, _ = orc.mlp_forward(enc)
        dW, d_enc = orc.mlp_backward(enc, hidden, np.ascontiguousarray(dy.view(np.uint16)))
        gx_ref = orc.grid_input_gradient(x, d_enc)
        a = gx.cpu().numpy()
        assert np.isfinite(a).all() and np.abs(gx_ref).max() > 0
        assert rae(a, gx_ref, 99.0) < 1e-2, interp  # tests/test_common.h:216
        g_ref = orc.backward_from_dy(x, dy.view(np.uint16))
        assert rae(ob.half_bits_to_float(f16(gp)), g_ref, 99.9) < 1.2e-2


def test_torch_layers_deliver_input_gradients(torch_cuda):
    """tcnn_b200.torch_modules: Encoding / NetworkWithInputEncoding return d(loss)/d(input) when the input requires grad
    (modules.py:153-160), equal to the native backward; Network (Identity encoding) evaluates."""
    torch = torch_cuda
    import tcnn_b200.torch_modules as tcnn

    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "configs", "hash3d_small.json")))
    enc = tcnn.Encoding(3, {**cfg["encoding"], "n_features_per_level": 4, "n_levels": 8})
    with torch.no_grad():
        enc.params.mul_(500.0)
    x = torch.rand(700, 3, device="cuda", requires_grad=True)
    y = enc(x)
    assert y.shape == (700, 32) and y.dtype == torch.float16
    (y.float() ** 2).sum().backward()
    assert x.grad is not None and x.grad.shape == (700, 3) and torch.isfinite(x.grad).all() and float(x.grad.abs().max()) > 0
    assert enc.params.grad is not None and float(enc.params.grad.abs().max()) > 0
    # what autograd delivers is the native backward's result (values are checked against the oracle in test_encoding_matches_oracle):
    # pad to 768 rows, dL/dy = 2 y scaled by the loss scale, native backward, unscale
    xp = torch.nn.functional.pad(x.detach(), [0, 0, 0, 768 - 700]).contiguous()
    p16 = enc.params.detach().half().contiguous()
    yp = enc.native_tcnn_module.fwd(xp, p16)
    dy = torch.zeros_like(yp)
    dy[:700] = (2 * yp[:700].float() * enc.loss_scale).half()
    _, gx = enc.native_tcnn_module.bwd(xp, p16, dy, want_params=False, want_input=True)
    assert torch.allclose(x.grad, gx[:700] / enc.loss_scale, rtol=1e-3, atol=1e-6 * float(x.grad.abs().max()))

    model = tcnn.NetworkWithInputEncoding(3, 3, cfg["encoding"], cfg["network"])
    x2 = torch.rand(512, 3, device="cuda", requires_grad=True)
    model(x2).float().sum().backward()
    assert x2.grad is not None and torch.isfinite(x2.grad).all()
    # the stand-alone network (tinycudann.Network): autograd w.r.t. parameters and inputs, against torch fp32 on the same fp16 weights
    net = tcnn.Network(3, 3, {"otype": "CutlassMLP", "n_neurons": 64, "n_hidden_layers": 2})
    x3 = torch.rand(300, 3, device="cuda", requires_grad=True)
    out = net(x3)
    assert out.shape == (300, 3) and out.dtype == torch.float16 and torch.isfinite(out).all()
    out.float().square().sum().backward()
    w = net.params.detach().half().float()
    W0, W1, Wo = w[: 64 * 16].view(64, 16).clone().requires_grad_(True), w[1024 : 1024 + 4096].view(64, 64).clone().requires_grad_(True), w[5120:].view(16, 64).clone().requires_grad_(True)
    xr = x3.detach().clone().requires_grad_(True)
    xin = torch.cat([xr.half().float(), torch.ones(300, 13, device="cuda")], 1)
    ref = (torch.relu(torch.relu(xin @ W0.T) @ W1.T) @ Wo.T)[:, :3]
    ref.square().sum().backward()
    want = torch.cat([W0.grad.flatten(), W1.grad.flatten(), Wo.grad.flatten()])
    assert torch.allclose(out.float(), ref, rtol=2e-2, atol=2e-3)
    assert float((net.params.grad - want).abs().mean()) < 2e-2 * float(want.abs().mean())
    assert float((x3.grad - xr.grad).abs().mean()) < 3e-2 * float(xr.grad.abs().mean())


@pytest.mark.parametrize("n_in,F", [(2, 2), (3, 4), (3, 8)])
def test_replicated_scatter_of_the_coarse_levels_matches_exact_sums(torch_cuda, n_in, F):
    """From 16 384 samples on, the coarse levels -- thousands of reductions per table entry -- scatter into private copies that a
    second kernel sums (grid_kernels.h plan_grid_scatter). Same addends, different order: compared with the oracle's exact sums, and
    a second call must give the same table (the copies are re-armed to zero)."""
    torch = torch_cuda
    import tcnn_b200

    L = 8
    enc_cfg = {"otype": "HashGrid", "n_levels": L, "n_features_per_level": F, "log2_hashmap_size": 14, "base_resolution": 8, "per_level_scale": 1.5}
    enc = tcnn_b200.Encoding(n_in, enc_cfg)
    levels = enc.grid_levels()
    orc = ob.OracleModel(n_in, 1, {"encoding": enc_cfg, "network": {"n_neurons": 16, "n_hidden_layers": 1}}, scales=levels["scales"])
    p16 = enc.initial_params(seed=9, scale=50.0).to(torch.float16).contiguous()
    B = 32768
    x = make_x(n_in, B)
    xd = torch.from_numpy(x).cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    dy = (torch.randn(B, L * F, device="cuda", generator=g) * 0.01).to(torch.float16).contiguous()
    gp, _ = enc.bwd(xd, p16, dy)
    gp2, _ = enc.bwd(xd, p16, dy)
    torch.cuda.synchronize()
    dy_soa = np.zeros((orc.grid.padded_width, B), np.uint16)
    dy_soa[: L * F] = f16(dy).T
    g_ref = orc.grid_backward(x, dy_soa)
    g_dev = ob.half_bits_to_float(f16(gp)).astype(np.float64)
    # coarse levels: sums of ~1 000 addends each; fp16 partial sums per copy + fp32 across copies are closer to the exact sums than
    # one fp16 accumulator would be
    assert rae(g_dev, g_ref, 99.9) < 1.2e-2
    assert ((g_dev != 0) != (g_ref != 0)).mean() < 2e-3
    first = levels["offsets"][1] * F  # level 0 alone
    assert rae(g_dev[:first], g_ref[:first]) < 1.2e-2
    assert rae(ob.half_bits_to_float(f16(gp2)), g_dev, 99.9) < 5e-3
