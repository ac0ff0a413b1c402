"""bench_configs.py -- training_step / inference throughput of the secondary configurations (BASELINE.json configs[0], configs[2], and the
n_features_per_level / 4-D variants of the general path), this library and -- with --reference -- the unmodified reference
(oracle/_ref/ref_harness bench, offline and JIT-fused) on the SAME batches: a pool of 4 drawn from pcg32{1337}, closed-form targets.

    python scripts/bench_configs.py [--configs image_w128,headline,...] [--steps 50] [--warmup 10] [--reference]

One JSON line per configuration. Device-resident inputs, CUDA events around the timed steps (the per-kernel split comes from the
library's own profiling events: binning / forward+backward / optimizer).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-cuda-nn_b200"))
sys.path.insert(0, ROOT)

OPT = {"otype": "Adam", "learning_rate": 1e-2, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6}


def hashgrid(**kw):
    enc = {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16, "per_level_scale": 1.5}
    enc.update(kw)
    return enc


def mlp(width, hidden, otype="FullyFusedMLP"):
    return {"otype": otype, "activation": "ReLU", "output_activation": "None", "n_neurons": width, "n_hidden_layers": hidden}


CONFIGS = {
    # name: (n_in, n_out, batch, config, which BASELINE.json row)
    "headline": (3, 3, 1 << 18, {"encoding": hashgrid(), "network": mlp(64, 2)}, "configs[1] (fused kernel)"),
    "image_w128": (2, 3, 1 << 18, {"encoding": hashgrid(log2_hashmap_size=15), "network": mlp(128, 4)}, "configs[2]: mlp_learning_an_image with 128 x 4 (general path)"),
    "image_w128_t19": (2, 3, 1 << 18, {"encoding": hashgrid(), "network": mlp(128, 4)}, "configs[2] with a 2^19 table (general path)"),
    "image_w64": (2, 3, 1 << 18, {"encoding": hashgrid(log2_hashmap_size=15), "network": mlp(64, 2)}, "samples/mlp_learning_an_image.cu with data/config_hash.json (fused kernel)"),
    "identity_cutlass": (3, 3, 1 << 16, {"encoding": {"otype": "Identity"}, "network": mlp(64, 2, "CutlassMLP")}, "configs[0] (fused kernel)"),
    "hash3d_w128": (3, 3, 1 << 18, {"encoding": hashgrid(), "network": mlp(128, 2)}, "headline encoding + 128 x 2 (general path)"),
    "f4_l8": (3, 3, 1 << 18, {"encoding": hashgrid(n_levels=8, n_features_per_level=4), "network": mlp(64, 2)}, "n_features_per_level = 4 (general path)"),
    # BASELINE.json configs[4], training half: the bare network behind an Identity encoding (the reference's bench_mlp_ours only times inference)
    "mlp_w128_h4": (128, 128, 1 << 18, {"encoding": {"otype": "Identity"}, "network": mlp(128, 4)}, "configs[4] training: 128 -> 128 x 4 -> 128 (general path)"),
    "mlp_w64_h4": (64, 64, 1 << 18, {"encoding": {"otype": "Identity"}, "network": mlp(64, 4)}, "configs[4] training: 64 -> 64 x 4 -> 64 (general path)"),
    "d4": (4, 3, 1 << 18, {"encoding": hashgrid(), "network": mlp(64, 2)}, "4-D inputs (general path)"),
}
for _c in CONFIGS.values():
    _c[3]["loss"] = {"otype": "RelativeL2"}
    _c[3]["optimizer"] = OPT


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="headline,image_w128,image_w64,identity_cutlass,hash3d_w128,f4_l8,d4")  # mlp_*: the harness builds 128-D targets on one host thread (minutes)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--reference", action="store_true")
    args = ap.parse_args()
    import torch

    import tcnn_b200
    from bench import closed_form_targets

    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    for name in args.configs.split(","):
        n_in, n_out, B, cfg, what = CONFIGS[name]
        model = tcnn_b200.create_from_config(n_in, n_out, cfg)
        trainer = model.trainer
        rng = tcnn_b200.Pcg32(1337)
        xs, ys = [], []
        for _ in range(4):
            x = tcnn_b200.generate_random_uniform(rng, B * n_in).view(B, n_in)
            xs.append(x)
            ys.append(closed_form_targets(torch, x, n_out))
        for i in range(args.warmup):
            trainer.training_step(xs[i % 4], ys[i % 4])
        torch.cuda.synchronize()
        model.set_profiling(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            trainer.training_step(xs[i % 4], ys[i % 4])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        prof = model.read_profile()
        model.set_profiling(False)
        trainer.training_step(xs[(args.warmup + args.steps) % 4], ys[(args.warmup + args.steps) % 4])
        final_loss = trainer.loss()
        for i in range(5):
            model.network.inference(xs[i % 4])
        torch.cuda.synchronize()
        e0.record()
        for i in range(args.steps):
            model.network.inference(xs[i % 4])
        e1.record()
        torch.cuda.synchronize()
        inf_ms = e0.elapsed_time(e1) / args.steps
        line = {"config": name, "what": what, "n_input_dims": n_in, "n_output_dims": n_out, "batch": B, "n_params": model.n_params, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "kernel_split_ms": prof, "loss_after_steps": final_loss,
                "inference_ms": inf_ms, "inference_samples_per_s": B / (inf_ms * 1e-3)}
        if args.reference and os.path.exists(harness):
            ref = {}
            with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
                json.dump(cfg, f)
            for mode, jit in (("fully_fused", 0), ("fully_fused_jit", 1)):
                for what_run, inf in (("training", 0), ("inference", 1)):
                    cmd = [harness, "bench", f.name, str(n_in), str(n_out), str(B), str(args.steps), str(args.warmup), str(jit), str(inf)]
                    try:
                        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
                        js = [l for l in out.stdout.splitlines() if l.startswith("{")]
                        d = json.loads(js[-1]) if js else {"error": (out.stderr or out.stdout)[-300:]}
                        ref[f"{mode}_{what_run}"] = {k: d[k] for k in ("ms_per_step", "samples_per_s", "loss_after_steps", "error") if k in d}
                    except Exception as e:  # noqa: BLE001
                        ref[f"{mode}_{what_run}"] = {"error": repr(e)}
            os.unlink(f.name)
            line["reference"] = ref
            best_t = max([v["samples_per_s"] for k, v in ref.items() if k.endswith("training") and "samples_per_s" in v], default=None)
            best_i = max([v["samples_per_s"] for k, v in ref.items() if k.endswith("inference") and "samples_per_s" in v], default=None)
            line["vs_reference_training"] = line["samples_per_s"] / best_t if best_t else None
            line["vs_reference_inference"] = line["inference_samples_per_s"] / best_i if best_i else None
        print(json.dumps(line), flush=True)
        del model, trainer, xs, ys
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
