"""The C++ source-compatibility headers (include/tiny-cuda-nn/*.h): an application written against tiny-cuda-nn's names
(tests/cpp/shim_sample.cu, modelled on the reference's samples/mlp_learning_an_image.cu) compiles against them, links
libtcnn_b200.so and trains. The binary is built by `__graft_entry__.build()` (nvcc cross-compiles without a GPU)."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "shim_sample")


def test_shim_headers_exist_and_cite_the_reference():
    inc = os.path.join(ROOT, "include", "tiny-cuda-nn")
    for h in ("common.h", "config.h", "gpu_matrix.h", "gpu_memory.h", "random.h", "trainer.h", "network_with_input_encoding.h", "loss.h", "optimizer.h"):
        text = open(os.path.join(inc, h)).read()
        assert "#pragma once" in text
    assert "trainer.h:254-357" in open(os.path.join(inc, "config.h")).read()


@pytest.mark.gpu
def test_shim_sample_trains(torch_cuda):
    if not os.path.exists(BIN):
        pytest.skip("tests/cpp/shim_sample has not been built (python __graft_entry__.py build)")
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "tiny-cuda-nn_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([BIN, "300"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    # HashGrid(2-D, L=16, F=2, T=2^15, base 16, s=1.5): 6 dense + 10 hashed levels = 354 184 entries + 64x32+64x64+16x64 weights
    assert r["n_params"] == 354184 * 2 + 7168 and r["padded_output_width"] == 16
    assert r["last_loss"] < 0.05 * r["first_loss"], r
    assert r["inference_mse"] < 5e-3, r
    assert np.isfinite(r["create_from_config_loss"]) and abs(r["create_from_config_loss"] - r["first_loss"]) / r["first_loss"] < 0.5
    # default_rng_t{1337} + generate_random_uniform == the reference's generator (random.h:40-69), via the oracle restatement
    expect = ob.generate_random_uniform(ob.default_rng(1337), 8)
    assert np.array_equal(np.array(r["first_uniform"], np.float32), expect)


def test_shim_sample_compiles_without_a_gpu():
    """nvcc cross-compiles the application against the compatibility headers (compile only, no link, no GPU)."""
    import shutil
    import sys
    import tempfile

    sys.path.insert(0, ROOT)
    from __graft_entry__ import json_include_dir

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    inc = json_include_dir()
    if not os.path.exists(nvcc) or inc is None:
        pytest.skip("nvcc or nlohmann/json.hpp not available")
    with tempfile.TemporaryDirectory() as tmp:
        out = subprocess.run([nvcc, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"), "-I", inc, "-c",
                              os.path.join(ROOT, "tests", "cpp", "shim_sample.cu"), "-o", os.path.join(tmp, "shim_sample.o")], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
