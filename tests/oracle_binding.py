"""ctypes binding of oracle/liboracle_cpu.so (the CPU restatement of the reference hot path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs. The product package (tiny-cuda-nn_b200/) never imports this.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "liboracle_cpu.so")

GRID_HASH, GRID_DENSE, GRID_TILED = 0, 1, 2
INTERP_NEAREST, INTERP_LINEAR, INTERP_SMOOTHSTEP = 0, 1, 2
ACT = {"relu": 0, "leakyrelu": 1, "silu": 2, "exponential": 3, "sine": 4, "sigmoid": 5, "squareplus": 6, "softplus": 7, "tanh": 8, "none": 9}
LOSS_L2, LOSS_RELATIVE_L2, LOSS_L1, LOSS_RELATIVE_L1, LOSS_MAPE, LOSS_SMAPE, LOSS_RELATIVE_L2_LUMINANCE, LOSS_CROSS_ENTROPY, LOSS_VARIANCE_IS = 0, 1, 2, 3, 4, 5, 6, 7, 8
ACCUM_FP32, ACCUM_FP16_K16 = 0, 1


class Grid(ctypes.Structure):
    _fields_ = [
        ("n_pos_dims", ctypes.c_uint32), ("n_levels", ctypes.c_uint32), ("n_features_per_level", ctypes.c_uint32),
        ("log2_hashmap_size", ctypes.c_uint32), ("base_resolution", ctypes.c_uint32), ("per_level_scale", ctypes.c_float),
        ("grid_type", ctypes.c_uint32), ("interpolation", ctypes.c_uint32), ("padded_width", ctypes.c_uint32),
        ("offsets", ctypes.c_uint32 * 129), ("scales", ctypes.c_float * 128), ("resolutions", ctypes.c_uint32 * 128),
        ("n_params", ctypes.c_uint32),
    ]


class Mlp(ctypes.Structure):
    _fields_ = [
        ("in_width", ctypes.c_uint32), ("width", ctypes.c_uint32), ("n_hidden_layers", ctypes.c_uint32), ("out_width", ctypes.c_uint32),
        ("padded_out_width", ctypes.c_uint32), ("activation", ctypes.c_uint32), ("output_activation", ctypes.c_uint32), ("n_params", ctypes.c_uint32),
    ]


class Adam(ctypes.Structure):
    _fields_ = [
        ("learning_rate", ctypes.c_float), ("beta1", ctypes.c_float), ("beta2", ctypes.c_float), ("epsilon", ctypes.c_float), ("l2_reg", ctypes.c_float),
        ("relative_decay", ctypes.c_float), ("absolute_decay", ctypes.c_float), ("clipping_magnitude", ctypes.c_float),
        ("gradient_clipping_magnitude", ctypes.c_float), ("non_matrix_learning_rate_factor", ctypes.c_float), ("non_matrix_l2_reg", ctypes.c_float),
        ("adabound", ctypes.c_int), ("optimize_matrix_params", ctypes.c_int), ("optimize_non_matrix_params", ctypes.c_int),
        ("skip_zero_grad_non_matrix_params", ctypes.c_int),
    ]


class Pcg32(ctypes.Structure):
    _fields_ = [("state", ctypes.c_uint64), ("inc", ctypes.c_uint64)]


class ModelDesc(ctypes.Structure):
    _fields_ = [
        ("grid", ctypes.POINTER(Grid)), ("mlp", ctypes.POINTER(Mlp)), ("adam", ctypes.POINTER(Adam)),
        ("loss_type", ctypes.c_int), ("accum_mode", ctypes.c_int), ("loss_scale", ctypes.c_float),
    ]


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "cpu"])


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = ctypes.CDLL(LIB)
        _lib.orc_pcg32_next_float.restype = ctypes.c_float
        _lib.orc_pcg32_next_uint.restype = ctypes.c_uint32
        _lib.orc_training_step.restype = ctypes.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def adam_from_config(opt):
    a = Adam(1e-3, 0.9, 0.999, 1e-8, 1e-8, 0, 0, 0, 0, 1.0, 0.0, 0, 1, 1, 1)
    for k in ("learning_rate", "beta1", "beta2", "epsilon", "l2_reg", "relative_decay", "absolute_decay", "clipping_magnitude",
              "gradient_clipping_magnitude", "non_matrix_learning_rate_factor", "non_matrix_l2_reg"):
        if k in opt:
            setattr(a, k, float(opt[k]))
    for k in ("adabound", "optimize_matrix_params", "optimize_non_matrix_params", "skip_zero_grad_non_matrix_params"):
        if k in opt:
            setattr(a, k, int(bool(opt[k])))
    return a


class OracleModel:
    """The reference's TrainableModel restated on the CPU: same JSON config, same initialisation, same step."""

    def __init__(self, n_in, n_out, config, seed=1337, accum_mode=ACCUM_FP32, scales=None):
        lib = load()
        self.lib = lib
        enc = config.get("encoding", {})
        net = config.get("network", {})
        otype = enc.get("otype", "HashGrid").lower()
        default_type = {"tiledgrid": "tiled", "densegrid": "dense"}.get(otype, "hash")
        gtype = {"hash": GRID_HASH, "dense": GRID_DENSE, "tiled": GRID_TILED}[enc.get("type", default_type).lower()]
        F = enc.get("n_features_per_level", 2)
        L = enc.get("n_levels", 16)
        base = enc.get("base_resolution", 16)
        default_scale = float(np.exp(np.log(np.float32(256.0) / np.float32(base)) / (L - 1))) if gtype == GRID_DENSE else 2.0
        interp = {"nearest": INTERP_NEAREST, "linear": INTERP_LINEAR, "smoothstep": INTERP_SMOOTHSTEP}[enc.get("interpolation", "Linear").lower()]
        self.grid = Grid(n_in, L, F, enc.get("log2_hashmap_size", 19), base, enc.get("per_level_scale", default_scale), gtype, interp, 0)
        assert lib.orc_grid_setup(ctypes.byref(self.grid)) == 0
        if scales is not None:
            for i, s in enumerate(scales):
                self.grid.scales[i] = s
        self.mlp = Mlp(self.grid.padded_width, net.get("n_neurons", 128), net.get("n_hidden_layers", 5), n_out, 0,
                       ACT[net.get("activation", "ReLU").lower()], ACT[net.get("output_activation", "None").lower()], 0)
        assert lib.orc_mlp_setup(ctypes.byref(self.mlp)) == 0
        self.adam = adam_from_config(config.get("optimizer", {}))
        self.loss_type = {"l2": LOSS_L2, "relativel2": LOSS_RELATIVE_L2, "l1": LOSS_L1, "relativel1": LOSS_RELATIVE_L1, "mape": LOSS_MAPE, "smape": LOSS_SMAPE, "relativel2luminance": LOSS_RELATIVE_L2_LUMINANCE, "crossentropy": LOSS_CROSS_ENTROPY, "variance": LOSS_VARIANCE_IS}[config.get("loss", {}).get("otype", "RelativeL2").lower()]
        self.desc = ModelDesc(ctypes.pointer(self.grid), ctypes.pointer(self.mlp), ctypes.pointer(self.adam), self.loss_type, accum_mode, 128.0)
        self.n_in, self.n_out = n_in, n_out
        self.n_mlp = self.mlp.n_params
        self.n_params = self.mlp.n_params + self.grid.n_params
        # Trainer ctor + initialize_params (trainer.h:51-87)
        self.rng = Pcg32()
        lib.orc_trainer_rng(ctypes.c_uint32(seed), ctypes.byref(self.rng))
        self.params_fp32 = np.zeros(self.n_params, np.float32)
        lib.orc_initialize_params(ctypes.byref(self.grid), ctypes.byref(self.mlp), ctypes.byref(self.rng), _p(self.params_fp32))
        self.params_fp16 = np.zeros(self.n_params, np.uint16)
        lib.orc_cast_to_half(ctypes.c_uint64(self.n_params), _p(self.params_fp32), _p(self.params_fp16))
        self.grads_fp16 = np.zeros(self.n_params, np.uint16)
        self.m1 = np.zeros(self.n_params, np.float32)
        self.m2 = np.zeros(self.n_params, np.float32)
        self.steps = np.zeros(self.n_params, np.uint32)

    def set_params_full_precision(self, p):
        self.params_fp32[:] = p
        self.lib.orc_cast_to_half(ctypes.c_uint64(self.n_params), _p(self.params_fp32), _p(self.params_fp16))

    def encode(self, x, want_indices=False):
        B = x.shape[0]
        enc = np.zeros((self.grid.padded_width, B), np.uint16)
        idx = np.zeros((B, self.grid.n_levels, 1 << self.n_in), np.uint32) if want_indices else None
        self.lib.orc_grid_forward(ctypes.byref(self.grid), B, _p(x), _p(self.params_fp16[self.n_mlp:].copy()), _p(enc), _p(idx))
        return (enc, idx) if want_indices else enc

    def mlp_forward(self, enc_soa):
        B = enc_soa.shape[1]
        hidden = np.zeros((self.mlp.n_hidden_layers, B, self.mlp.width), np.uint16)
        out = np.zeros((B, self.mlp.padded_out_width), np.uint16)
        self.lib.orc_mlp_forward(ctypes.byref(self.mlp), B, self.desc.accum_mode, _p(self.params_fp16), _p(enc_soa), _p(hidden), _p(out))
        return hidden, out

    def loss(self, out16, targets):
        B = out16.shape[0]
        values = np.zeros((B, self.mlp.padded_out_width), np.float32)
        grads = np.zeros((B, self.mlp.padded_out_width), np.uint16)
        self.lib.orc_loss(self.loss_type, B, self.mlp.padded_out_width, self.n_out, ctypes.c_float(128.0), _p(out16), _p(targets), _p(values), _p(grads))
        return values, grads

    def mlp_backward(self, enc_soa, hidden, dL_dout):
        B = enc_soa.shape[1]
        dW = np.zeros(self.n_mlp, np.float64)
        d_enc = np.zeros_like(enc_soa)
        self.lib.orc_mlp_backward(ctypes.byref(self.mlp), B, self.desc.accum_mode, _p(self.params_fp16), _p(enc_soa), _p(hidden), _p(dL_dout), _p(dW), _p(d_enc))
        return dW, d_enc

    def grid_backward(self, x, d_enc_soa):
        g = np.zeros(self.grid.n_params, np.float64)
        self.lib.orc_grid_backward(ctypes.byref(self.grid), x.shape[0], _p(x), _p(d_enc_soa), _p(g))
        return g

    def grid_input_gradient(self, x, d_enc_soa, grid_params_fp16=None):
        """dL/d(position) [B][D] fp32 from dL/d(encoded) SoA fp16 bits (grid.h:171-210 + 322-350)."""
        out = np.zeros((x.shape[0], self.n_in), np.float32)
        table = np.ascontiguousarray(self.params_fp16[self.n_mlp:] if grid_params_fp16 is None else grid_params_fp16)
        self.lib.orc_grid_input_gradient(ctypes.byref(self.grid), x.shape[0], _p(x), _p(table), _p(np.ascontiguousarray(d_enc_soa)), _p(out))
        return out

    def backward_from_dy(self, x, dL_dout_bits):
        """Module-tier backward (cpp_api.cu:105-117) restated from the stage functions: fp16-rounded gradients of all parameters
        for a caller-supplied dL/d(output) [B][padded_out] (fp16 bit patterns). Output activation None only."""
        enc = self.encode(x)
        hidden, _ = self.mlp_forward(enc)
        dW, d_enc = self.mlp_backward(enc, hidden, np.ascontiguousarray(dL_dout_bits))
        dG = self.grid_backward(x, d_enc)
        return np.concatenate([dW, dG]).astype(np.float16).astype(np.float32)

    def training_step(self, x, y, run_optimizer=True, want_loss_values=False):
        B = x.shape[0]
        lv = np.zeros((B, self.n_out), np.float32) if want_loss_values else None
        s = self.lib.orc_training_step(ctypes.byref(self.desc), B, _p(x), _p(y), _p(self.params_fp32), _p(self.params_fp16), _p(self.grads_fp16),
                                       _p(self.m1), _p(self.m2), _p(self.steps), int(run_optimizer), _p(lv))
        return (s, lv) if want_loss_values else s

    def inference(self, x):
        out = np.zeros((x.shape[0], self.n_out), np.float32)
        self.lib.orc_inference(ctypes.byref(self.desc), x.shape[0], _p(x), _p(self.params_fp16), _p(out))
        return out


def generate_random_uniform(rng, n, lo=0.0, hi=1.0):
    """random.h:56-69 on a Pcg32 state (advanced in place)."""
    out = np.zeros(n, np.float32)
    load().orc_generate_random_uniform(ctypes.byref(rng), ctypes.c_uint64(n), _p(out), ctypes.c_float(lo), ctypes.c_float(hi))
    return out


def default_rng(seed=1337):
    """default_rng_t rng{seed} == pcg32(seed, 1) (samples/mlp_learning_an_image.cu:222)."""
    rng = Pcg32()
    load().orc_pcg32_seed(ctypes.byref(rng), ctypes.c_uint64(seed), ctypes.c_uint64(1))
    return rng


def make_targets(x, n_out):
    """Closed-form target field of oracle/ref_harness.cu::make_targets, evaluated in fp32 in the same order."""
    B, n_in = x.shape
    y = np.zeros((B, n_out), np.float32)
    for c in range(n_out):
        phase = np.zeros(B, np.float32)
        for d in range(n_in):
            phase = (phase + (x[:, d] * np.float32(c + 1 + d)) / np.float32(1 << d)).astype(np.float32)
        y[:, c] = (np.float32(0.5) + np.float32(0.5) * np.sin((np.float32(6.2831853) * phase).astype(np.float32)).astype(np.float32)).astype(np.float32)
    return y


def half_bits_to_float(a):
    return a.view(np.float16).astype(np.float32)


class OracleShardTrainer:
    """Stand-in for the CUDA trainer with the interface tcnn_b200.dp.DataParallelTrainer drives, backed by the CPU oracle.
    Lets the data-parallel host logic (sharding, global loss normalisation, gradient all-reduce, replicated Adam) be
    tested with gloo on CPU. Gradient buffers are fp64 exact sums (the device all-reduces fp16 / fp32 partial sums)."""

    def __init__(self, n_in, n_out, config, seed=1337):
        import torch

        self.m = m = OracleModel(n_in, n_out, config, seed=seed)
        # padded whole-vector buffers like the CUDA trainer's (tcnnb_n_params_padded): the model's arrays become prefixes of them
        self.n_pad = (m.n_params + 511) // 512 * 512
        self._p16 = np.zeros(self.n_pad, np.uint16)
        self._p32 = np.zeros(self.n_pad, np.float32)
        self._p16[: m.n_params] = m.params_fp16
        self._p32[: m.n_params] = m.params_fp32
        m.params_fp16 = self._p16[: m.n_params]
        m.params_fp32 = self._p32[: m.n_params]
        self._grads = torch.zeros(self.n_pad, dtype=torch.float64)
        self.grad_sums = self._grads[: m.n_params]
        self._loss = 0.0
        load().orc_training_step_shard.restype = ctypes.c_double

    def training_step_shard(self, x, y, global_batch, run_optimizer=False):
        m = self.m
        xs, ys = np.ascontiguousarray(x.numpy()), np.ascontiguousarray(y.numpy())
        self._loss = m.lib.orc_training_step_shard(ctypes.byref(m.desc), xs.shape[0], int(global_batch), _p(xs), _p(ys), _p(m.params_fp32), _p(m.params_fp16),
                                                   _p(m.grads_fp16), ctypes.c_void_p(self.grad_sums.data_ptr()), _p(m.m1), _p(m.m2), _p(m.steps), 0, None)
        if run_optimizer:
            self.optimizer_step()

    def gradient_buffers(self):
        return [self.grad_sums]

    def sharded_buffers(self):
        import torch

        return {"grads": self._grads, "params": torch.from_numpy(self._p16.view(np.int16)), "masters": torch.from_numpy(self._p32),
                "n_params": self.m.n_params, "n_matrix": self.m.n_mlp}

    def finalize_gradients(self):
        pass  # one fp64 gradient vector already

    def optimizer_step(self, ranges=None):
        m = self.m
        m.grads_fp16[:] = self.grad_sums.numpy().astype(np.float16).view(np.uint16)
        for b, c in ranges if ranges is not None else [(0, m.n_params)]:
            if c == 0:
                continue
            n_matrix = max(0, min(m.n_mlp - b, c))
            sl = slice(b, b + c)
            m.lib.orc_adam_step(ctypes.byref(m.adam), ctypes.c_uint64(c), ctypes.c_uint64(n_matrix), ctypes.c_float(128.0), _p(m.params_fp32[sl]), _p(m.params_fp16[sl]),
                                _p(m.grads_fp16[sl]), _p(m.m1[sl]), _p(m.m2[sl]), _p(m.steps[sl]))

    def loss(self):
        return self._loss
