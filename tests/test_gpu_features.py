"""GPU parity of the parameter-free encodings and of Composite (encodings/{identity,frequency,triangle_wave,oneblob,
spherical_harmonics,composite}.h): csrc/feature_encodings.cu + grid_kernels.cu through the encoding tier (tcnnb_encoding_*) and through
create_from_config on the general path, against the numpy oracle (oracle/feature_encodings.py), the C oracle (grid, network, loss) and --
when the fixtures exist -- the vectors the unmodified reference wrote (tests/golden/composite_*.npz)."""
import ctypes
import json
import os
import sys

import numpy as np
import pytest

import oracle_binding as ob
from golden_util import GOLDEN, rae

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import feature_encodings as fe  # noqa: E402

pytestmark = pytest.mark.gpu

GRID3 = {"otype": "HashGrid", "n_levels": 8, "n_features_per_level": 2, "log2_hashmap_size": 14, "base_resolution": 8, "per_level_scale": 1.5}
ENCODINGS = {
    "identity": (5, {"otype": "Identity", "scale": 2.0, "offset": -0.5}),
    "frequency": (3, {"otype": "Frequency", "n_frequencies": 6}),
    "triangle_wave": (2, {"otype": "TriangleWave", "n_frequencies": 8}),
    "oneblob": (3, {"otype": "OneBlob", "n_bins": 16}),
    "sh1": (3, {"otype": "SphericalHarmonics", "degree": 1}),
    "sh4": (3, {"otype": "SphericalHarmonics", "degree": 4}),
    "sh8": (3, {"otype": "SphericalHarmonics", "degree": 8}),
    "nrc": (6, {"otype": "Composite", "nested": [{"n_dims_to_encode": 3, "otype": "TriangleWave", "n_frequencies": 4}, {"n_dims_to_encode": 2, "otype": "OneBlob", "n_bins": 4}, {"otype": "Identity"}]}),
    "grid_sh": (6, {"otype": "Composite", "nested": [dict(GRID3, n_dims_to_encode=3), {"n_dims_to_encode": 3, "otype": "SphericalHarmonics", "degree": 4}]}),
    "freq_grid4": (5, {"otype": "Composite", "nested": [{"n_dims_to_encode": 3, "otype": "Frequency", "n_frequencies": 1}, dict(GRID3, n_dims_to_encode=2, n_features_per_level=4, n_levels=2)]}),
}


def f16(t):
    return t.cpu().numpy().view(np.uint16)


def make_x(n_in, B, seed=1337):
    rng = ob.default_rng(seed)
    return ob.generate_random_uniform(rng, B * n_in).reshape(B, n_in)


def grid_oracle(sub_cfg, n_dims, scales):
    return ob.OracleModel(n_dims, 1, {"encoding": {k: v for k, v in sub_cfg.items() if k != "n_dims_to_encode"}, "network": {"n_neurons": 16, "n_hidden_layers": 1}}, scales=scales)


@pytest.mark.parametrize("name", list(ENCODINGS))
def test_encoding_tier_matches_the_oracles(torch_cuda, name):
    torch = torch_cuda
    import tcnn_b200

    n_in, cfg = ENCODINGS[name]
    enc = tcnn_b200.Encoding(n_in, cfg)
    B = 1024
    x = make_x(n_in, B)
    xd = torch.from_numpy(x).cuda()
    want, segs, width = fe.encode_plain(x, cfg, 0)
    assert enc.n_output_dims == width and enc.n_input_dims == n_in
    p16 = enc.initial_params(seed=9, scale=50.0).to(torch.float16).contiguous() if enc.n_params else torch.zeros(8, dtype=torch.float16, device="cuda")
    out = enc.fwd(xd, p16)
    g = torch.Generator(device="cuda").manual_seed(3)
    dy = (torch.randn(B, width, device="cuda", generator=g) * 0.1).to(torch.float16).contiguous()
    gp, gx = enc.bwd(xd, p16, dy, want_params=enc.n_params > 0, want_input=True)
    torch.cuda.synchronize()
    got = out.float().cpu().numpy()
    plain = ~np.isnan(want)
    # fp32 evaluation rounded to fp16; the reference's fast-math sin differs from libm's in the last bits: a few fp16 ulps of values <= ~3
    # (relative for the high-degree harmonics of non-unit directions, whose values reach ~50)
    assert (np.abs(got[plain] - want[plain]) <= 4e-3 * np.maximum(1.0, np.abs(want[plain]))).all(), np.abs(got[plain] - want[plain]).max()
    assert (got[plain].astype(np.float16) != want[plain].astype(np.float16)).mean() < 0.05

    # input gradients: analytic oracle per nested encoding; grids through the C oracle
    dyf = dy.float().cpu().numpy()
    gx_ref = np.zeros((B, n_in), np.float32)
    param_offset = 0
    levels = enc.grid_levels()
    for t, e, in_begin, n_dims, out_begin, n_out, n_pad, pad_first in segs:
        xs = x[:, in_begin : in_begin + n_dims]
        first = out_begin + (n_pad if pad_first else 0)
        dys = dyf[:, first : first + n_out]
        if t == "identity":
            gx_ref[:, in_begin : in_begin + n_dims] = dys * np.float32(e.get("scale", 1.0))
        elif t == "frequency":
            gx_ref[:, in_begin : in_begin + n_dims] = fe.frequency_input_gradient(xs, dys, e.get("n_frequencies", 12))
        elif t == "trianglewave":
            gx_ref[:, in_begin : in_begin + n_dims] = fe.triangle_wave_input_gradient(xs, dys, e.get("n_frequencies", 12))
        elif t == "oneblob":
            gx_ref[:, in_begin : in_begin + n_dims] = fe.oneblob_input_gradient(xs, dys, e.get("n_bins", 16))
        elif t == "sphericalharmonics":
            gx_ref[:, in_begin : in_begin + n_dims] = fe.spherical_harmonics_input_gradient(xs, dys, e.get("degree", 4))
        else:
            orc = grid_oracle(e, n_dims, levels["scales"])
            n_p = orc.grid.n_params
            orc.params_fp16[orc.n_mlp :] = f16(p16)[param_offset : param_offset + n_p]
            xs = np.ascontiguousarray(xs)
            enc_ref = orc.encode(xs)
            assert np.array_equal(f16(out)[:, out_begin : out_begin + n_out].T, enc_ref[:n_out]), "grid features inside the composite differ"
            assert (got[:, out_begin + n_out : out_begin + n_out + n_pad] == 0).all()
            dy_soa = np.zeros((orc.grid.padded_width, B), np.uint16)
            dy_soa[:n_out] = f16(dy)[:, out_begin : out_begin + n_out].T
            gx_ref[:, in_begin : in_begin + n_dims] = orc.grid_input_gradient(xs, dy_soa)
            g_ref = orc.grid_backward(xs, dy_soa)
            g_dev = ob.half_bits_to_float(f16(gp))[param_offset : param_offset + n_p]
            assert rae(g_dev, g_ref, 99.9) < 1.2e-2
            param_offset += n_p
    gx_dev = gx.cpu().numpy()
    scale = max(1.0, float(np.abs(gx_ref).max()))
    assert np.abs(gx_dev - gx_ref).max() <= 5e-3 * scale, (np.abs(gx_dev - gx_ref).max(), scale)


def net_cfg(width=64, hidden=2):
    return {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": width, "n_hidden_layers": hidden}


OPT = {"otype": "Adam", "learning_rate": 1e-2, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6}


@pytest.mark.parametrize("name", ["nrc", "grid_sh", "frequency", "sh4"])
def test_training_step_with_composite_encodings_matches_the_oracles(torch_cuda, name):
    """create_from_config on the general path: encoded features (numpy oracle + C grid oracle) -> network forward / loss / backward (C oracle)
    -> table gradients (C oracle), against the device's loss and gradients; then the loss goes down over ten steps."""
    torch = torch_cuda
    import tcnn_b200

    n_in, enc_cfg = ENCODINGS[name]
    n_out, B = 3, 512
    cfg = {"loss": {"otype": "RelativeL2"}, "optimizer": OPT, "encoding": enc_cfg, "network": net_cfg()}
    model = tcnn_b200.create_from_config(n_in, n_out, cfg)
    x = make_x(n_in, B)
    y = ob.make_targets(x[:, :3].copy(), n_out)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    model.trainer.training_step(xd, yd, run_optimizer=False)
    loss_dev = model.trainer.loss()
    grads = model.trainer.param_gradients().float().cpu().numpy()
    p16 = f16(model.trainer.params())

    enc, segs, width = fe.encode_plain(x, enc_cfg, 16)
    assert width == model.encoded_width
    n_mlp = model.n_mlp_params
    grid_parts, off = [], n_mlp
    for t, e, in_begin, n_dims, out_begin, n_o, n_pad, pad_first in segs:
        if t in ("grid", "hashgrid"):
            scales = tcnn_b200.Encoding(n_dims, {k: v for k, v in e.items() if k != "n_dims_to_encode"}).grid_levels()["scales"]
            orc = grid_oracle(e, n_dims, scales)
            orc.params_fp16[orc.n_mlp :] = p16[off : off + orc.grid.n_params]
            xs = np.ascontiguousarray(x[:, in_begin : in_begin + n_dims])
            enc[:, out_begin : out_begin + n_o] = ob.half_bits_to_float(orc.encode(xs)[:n_o].T)
            grid_parts.append((orc, xs, out_begin, n_o, off))
            off += orc.grid.n_params
    assert off == model.n_params and not np.isnan(enc).any()
    lib = ob.load()
    mlp = ob.Mlp(width, 64, 2, 16, 0, ob.ACT["relu"], ob.ACT["none"], 0)
    assert lib.orc_mlp_setup(ctypes.byref(mlp)) == 0 and mlp.n_params == n_mlp
    enc_soa = np.ascontiguousarray(enc.astype(np.float16).view(np.uint16).T)
    hidden = np.zeros((2, B, 64), np.uint16)
    out = np.zeros((B, 16), np.uint16)
    w16 = np.ascontiguousarray(p16[:n_mlp])
    lib.orc_mlp_forward(ctypes.byref(mlp), B, ob.ACCUM_FP32, ob._p(w16), ob._p(enc_soa), ob._p(hidden), ob._p(out))
    values = np.zeros((B, 16), np.float32)
    dy = np.zeros((B, 16), np.uint16)
    lib.orc_loss(ob.LOSS_RELATIVE_L2, B, 16, n_out, ctypes.c_float(128.0), ob._p(out), ob._p(y), ob._p(values), ob._p(dy))
    assert abs(loss_dev - values.sum()) <= 2e-3 * values.sum(), (loss_dev, values.sum())
    dW = np.zeros(n_mlp, np.float64)
    d_enc = np.zeros_like(enc_soa)
    lib.orc_mlp_backward(ctypes.byref(mlp), B, ob.ACCUM_FP32, ob._p(w16), ob._p(enc_soa), ob._p(hidden), ob._p(dy), ob._p(dW), ob._p(d_enc))
    assert rae(grads[:n_mlp], dW.astype(np.float16).astype(np.float32), 99.9) < 2e-2
    for orc, xs, out_begin, n_o, off in grid_parts:
        dy_soa = np.zeros((orc.grid.padded_width, B), np.uint16)
        dy_soa[:n_o] = d_enc[out_begin : out_begin + n_o]
        g_ref = orc.grid_backward(xs, dy_soa)
        g_dev = grads[off : off + orc.grid.n_params]
        assert ((g_dev != 0) != (g_ref != 0)).mean() < 5e-3
        assert rae(g_dev, g_ref, 99.9) < 2e-2

    losses = []
    for _ in range(10):
        model.trainer.training_step(xd, yd)
        losses.append(model.trainer.loss())
    assert losses[-1] < 0.7 * losses[0], losses
    inf = model.network.inference(xd)
    assert torch.isfinite(inf).all()


@pytest.mark.parametrize("case", ["composite_nrc", "composite_grid_sh", "frequency_top"])
def test_against_reference_golden_vectors_with_composite_encodings(torch_cuda, case):
    path = os.path.join(GOLDEN, case + ".npz")
    if not os.path.exists(path):
        pytest.skip("fixture not generated yet (tests/golden/make_golden.sh on a GPU box)")
    torch = torch_cuda
    import tcnn_b200

    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    cfg = json.load(open(os.path.join(GOLDEN, "configs", case + ".json")))
    n_in, n_out, B = meta["n_in"], meta["n_out"], meta["batch"]
    model = tcnn_b200.create_from_config(n_in, n_out, cfg)
    x, y = z["x_f32"].reshape(B, n_in), z["y_f32"].reshape(B, n_out)
    xd, yd = torch.from_numpy(x.copy()).cuda(), torch.from_numpy(y.copy()).cuda()
    p0 = model.trainer.params_full_precision().cpu().numpy()
    assert np.array_equal(p0.view(np.uint32), z["params_init_f32"].view(np.uint32))  # same pcg32 streams through the nested encodings
    inf = model.network.inference(xd).cpu().numpy()
    assert rae(inf, z["inference_f32"].reshape(B, n_out), 99.0) < 1e-2
    model.trainer.training_step(xd, yd, run_optimizer=False)
    loss0 = model.trainer.loss()
    assert abs(loss0 - meta["losses"][0]) <= 2e-3 * meta["losses"][0]
    grads = ob.half_bits_to_float(f16(model.trainer.param_gradients()))
    ref = ob.half_bits_to_float(z["grads_step0_f16"])
    n_mlp = model.n_mlp_params
    assert rae(grads[:n_mlp], ref[:n_mlp], 99.9) < 2e-2
    if model.n_params > n_mlp:
        assert rae(grads[n_mlp:], ref[n_mlp:], 99.9) < 1.2e-2
        assert ((grads[n_mlp:] != 0) != (ref[n_mlp:] != 0)).mean() < 2e-3
    losses = [loss0]
    for _ in range(meta["n_steps"]):
        model.trainer.training_step(xd, yd)
        losses.append(model.trainer.loss())
    for mine, theirs in zip(losses[1:], meta["losses"][1:]):
        assert abs(mine - theirs) <= 3e-2 * abs(theirs), (losses, meta["losses"])
