"""The C-ABI library loads without a GPU and exports every symbol include/tcnn_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tcnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tcnnb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import tcnn_b200

    lib = ctypes.CDLL(tcnn_b200.lib_path())
    declared = _declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    bound = {name for name, _, _ in tcnn_b200.ABI}
    assert set(declared) == bound, set(declared) ^ bound


def test_scalar_entry_points_without_gpu():
    import tcnn_b200

    lib = tcnn_b200.load()
    assert lib.tcnnb_abi_version() == 1
    assert lib.tcnnb_batch_size_granularity() == 256
    assert lib.tcnnb_default_loss_scale() == 128.0


def test_config_errors_are_reported_not_swallowed():
    """create_from_config validates the JSON before touching the device: bad configs fail with the reference's messages."""
    import torch

    import tcnn_b200

    if torch.cuda.is_available():
        pytest.skip("error-path test for the GPU-less container")
    with pytest.raises(tcnn_b200.TcnnError):
        tcnn_b200.create_from_config(3, 3, {"encoding": {"otype": "HashGrid"}, "network": {"otype": "FullyFusedMLP", "n_neurons": 64, "n_hidden_layers": 2}})


def test_json_reader_roundtrip():
    # the library's own JSON reader is exercised through hyperparams() on the GPU; here: the python side produces valid text
    import json

    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "configs", "headline.json")))
    assert cfg["encoding"]["log2_hashmap_size"] == 19 and cfg["network"]["n_neurons"] == 64


def test_python_pcg32_matches_the_oracle_generator_state():
    """tcnn_b200.Pcg32 (what bench.py seeds generate_random_uniform with) == pcg32{seed} of the reference (oracle restatement):
    same (state, inc) after seeding and after jumping ahead by the number of values a fill consumes."""
    import ctypes

    import oracle_binding as ob
    import tcnn_b200

    for seed in (1337, 1338, 42):
        a, b = tcnn_b200.Pcg32(seed), ob.default_rng(seed)
        assert (a.state, a.inc) == (b.state, b.inc)
        n = 3 * 1000 + 7
        ob.generate_random_uniform(b, n)  # advances the oracle generator by n
        a.advance(n)
        assert (a.state, a.inc) == (b.state, b.inc)
