"""PyTorch autograd layer on the module tier -- the role of tinycudann's `modules.py` (bindings/torch/tinycudann/modules.py:132-330).

    import tcnn_b200.torch_modules as tcnn
    model = tcnn.NetworkWithInputEncoding(n_input_dims=3, n_output_dims=3, encoding_config=enc, network_config=net)
    y = model(x)                      # x: [B, n_input_dims] float CUDA tensor -> [B, n_output_dims] fp16
    loss.backward(); torch_optimizer.step()   # model.params is a torch.nn.Parameter (fp32), as in tinycudann

Same conventions as the reference binding: parameters are an fp32 `torch.nn.Parameter` initialised by
`Module::initialize_params(seed)`, cast to fp16 for every call (modules.py:227-231); the batch is padded to the granularity of
256 (modules.py:222-226); gradients w.r.t. the output are multiplied by the loss scale (128 for fp16) before the native
backward pass and the parameter gradients divided by it afterwards (modules.py:166-171); the padded output columns are sliced
away (modules.py:233). Not (yet) supported: gradients w.r.t. the input and second-order terms -- an input that requires grad
raises instead of silently returning nothing.

This file is glue: the work happens in `tcnnb_module_forward` / `tcnnb_module_backward` (libtcnn_b200, fused sm_100a kernels).
"""
import torch

from . import Module as _NativeModule
from . import load

_BATCH_GRANULARITY = 256


class _ModuleFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, native, inputs, params, loss_scale):
        ctx.set_materialize_grads(False)
        output = native.fwd(inputs, params)
        ctx.save_for_backward(inputs, params, output)
        ctx.native = native
        ctx.loss_scale = loss_scale
        return output

    @staticmethod
    def backward(ctx, doutput):
        if doutput is None:
            return None, None, None, None
        inputs, params, output = ctx.saved_tensors
        with torch.no_grad():
            scaled = (doutput.to(torch.float32) * ctx.loss_scale).to(torch.float16).contiguous()
            grads = ctx.native.bwd(inputs, params, scaled, output=output)
            grads = (grads.to(torch.float32) / ctx.loss_scale).to(params.dtype)
        return None, None, grads, None


class NetworkWithInputEncoding(torch.nn.Module):
    """tinycudann.NetworkWithInputEncoding (modules.py:270-310) for the HashGrid + FullyFusedMLP path."""

    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed=1337):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_output_dims
        self.encoding_config = encoding_config
        self.network_config = network_config
        self.seed = seed
        self.native_tcnn_module = _NativeModule(n_input_dims, n_output_dims, encoding_config, network_config)
        self.dtype = torch.float16
        self.params = torch.nn.Parameter(self.native_tcnn_module.initial_params(seed), requires_grad=True)
        self.loss_scale = float(load().tcnnb_default_loss_scale())

    def forward(self, x):
        if x.requires_grad:
            raise NotImplementedError("tcnn_b200: gradients w.r.t. the input positions are not implemented")
        x = x.cuda() if not x.is_cuda else x
        batch = x.shape[0]
        padded = (batch + _BATCH_GRANULARITY - 1) // _BATCH_GRANULARITY * _BATCH_GRANULARITY
        if padded != batch:
            x = torch.nn.functional.pad(x, [0, 0, 0, padded - batch])
        out = _ModuleFunction.apply(self.native_tcnn_module, x.to(torch.float32).contiguous(), self.params.to(torch.float16).contiguous(), self.loss_scale)
        return out[:batch, : self.n_output_dims]

    def extra_repr(self):
        return f"n_input_dims={self.n_input_dims}, n_output_dims={self.n_output_dims}, seed={self.seed}, dtype={self.dtype}"
