"""examples/nerf_network.py -- the instant-ngp NeRF network (BASELINE.json configs[3]) assembled from two modules of this library, the way
instant-ngp's NerfNetwork assembles it from tiny-cuda-nn pieces and the way users of the PyTorch binding write it:

    position (3) --HashGrid(L=16, F=2, T=2^19)--> MLP 64 x 1 --> 16 features (feature 0 = log density)
    direction (3) --SphericalHarmonics(degree 4)--+
                                                  +--> Composite [SH | Identity(16 features)] --> MLP 64 x 2 --> rgb (3)

The density half runs on the fused kernel (fused_ws.cu), the colour half on the general path (feature_encodings.cu + the stand-alone
network kernels); autograd chains them through the input gradient of the colour module (tcnnb_module_backward, dL_dinput).

    python examples/nerf_network.py [batch] [steps]     # random rays x samples, synthetic targets: throughput of fwd + bwd + Adam
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-cuda-nn_b200"))
import torch  # noqa: E402

import tcnn_b200.torch_modules as tcnn  # noqa: E402

DENSITY_ENCODING = {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16, "per_level_scale": 1.5}
DENSITY_NETWORK = {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 1}
COLOR_ENCODING = {"otype": "Composite", "nested": [{"n_dims_to_encode": 3, "otype": "SphericalHarmonics", "degree": 4}, {"otype": "Identity", "n_dims_to_encode": 16}]}
COLOR_NETWORK = {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 2}


class NerfNetwork(torch.nn.Module):
    def __init__(self, seed=1337):
        super().__init__()
        self.density = tcnn.NetworkWithInputEncoding(3, 16, DENSITY_ENCODING, DENSITY_NETWORK, seed=seed)
        self.color = tcnn.NetworkWithInputEncoding(3 + 16, 3, COLOR_ENCODING, COLOR_NETWORK, seed=seed + 1)

    def forward(self, positions, directions):
        """positions, directions in [0, 1]^3 (directions as (d + 1) / 2, the convention of the SphericalHarmonics encoding) ->
        (rgb [n, 3] in [0, 1], density [n] >= 0)"""
        features = self.density(positions)                      # fp16 [n, 16]
        rgb_raw = self.color(torch.cat([directions.to(torch.float32), features.to(torch.float32)], dim=1))
        return torch.sigmoid(rgb_raw.float()), torch.exp(features[:, 0].float().clamp(max=10.0))


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    model = NerfNetwork()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, eps=1e-15)
    pos = torch.rand(batch, 3, device="cuda")
    direction = torch.nn.functional.normalize(torch.randn(batch, 3, device="cuda"), dim=1) * 0.5 + 0.5
    target_rgb = 0.5 + 0.5 * torch.sin(6.28 * (pos * torch.tensor([1.0, 2.0, 3.0], device="cuda")).cumsum(1))
    target_sigma = pos.sum(1)

    def step():
        opt.zero_grad(set_to_none=True)
        rgb, sigma = model(pos, direction)
        loss = ((rgb - target_rgb) ** 2).mean() + 0.1 * ((sigma - target_sigma) ** 2).mean()
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        first = step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(f"batch {batch}: {dt * 1e3:.3f} ms per step (forward + backward + torch Adam over {sum(p.numel() for p in model.parameters())} parameters), "
          f"{batch / dt:.3e} samples/s, loss {float(first):.4f} -> {float(loss):.4f}")


if __name__ == "__main__":
    main()
