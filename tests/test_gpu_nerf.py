"""The instant-ngp NeRF network (BASELINE.json configs[3]) assembled from two modules of this library
(examples/nerf_network.py): HashGrid -> 64 x 1 -> 16 features on the fused kernel, Composite[SphericalHarmonics(4), Identity(16)] -> 64 x 2 -> rgb on
the general path, chained by autograd through the colour module's input gradient."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nerf_network_trains_through_both_modules(torch_cuda):
    torch = torch_cuda
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import nerf_network

    torch.manual_seed(0)
    model = nerf_network.NerfNetwork()
    assert model.density.params.numel() == 64 * 32 + 16 * 64 + 13026992  # W0 [64][32], W_out [16][64], the 2^19-entry table of the headline grid
    assert model.color.params.numel() == 64 * 32 + 64 * 64 + 16 * 64  # Composite[SH 16 | Identity 16] -> 64 x 2 -> 16 (padded) outputs, no table
    n = 4096 + 37  # not a multiple of 256: the binding pads
    pos = torch.rand(n, 3, device="cuda", requires_grad=True)
    direction = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=1) * 0.5 + 0.5
    rgb, sigma = model(pos, direction)
    assert rgb.shape == (n, 3) and sigma.shape == (n,) and torch.isfinite(rgb).all() and torch.isfinite(sigma).all()
    target = torch.rand(n, 3, device="cuda")
    # (summed, not averaged: with a mean over 12 k elements the density network's weight gradients -- products of 1e-4-sized encodings --
    # fall below fp16's subnormal range behind the binding's loss scale of 128, here as in the reference's binding)
    loss = ((rgb - target) ** 2).sum()
    loss.backward()
    # the colour loss reaches the hash table and the positions THROUGH the colour module's dL/d(input) (its 16 feature inputs)
    gd, gc = model.density.params.grad, model.color.params.grad
    assert gd is not None and gc is not None and torch.isfinite(gd).all() and torch.isfinite(gc).all()
    assert float(gc.abs().sum()) > 0 and float(gd[: 32 * 64 + 16 * 64].abs().sum()) > 0 and float(gd[3072:].abs().sum()) > 0
    assert pos.grad is not None and torch.isfinite(pos.grad).all() and float(pos.grad.abs().sum()) > 0
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, eps=1e-15)
    pos = pos.detach()
    losses = []
    for _ in range(30):
        opt.zero_grad(set_to_none=True)
        rgb, _ = model(pos, direction)
        loss = ((rgb - target) ** 2).sum() / 64
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.6 * losses[0], losses
