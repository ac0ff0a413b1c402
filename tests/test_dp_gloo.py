"""Data-parallel host logic on CPU: two gloo ranks, each holding a replica, must reproduce the single-process step.
(The CUDA trainer is swapped for an oracle-backed stand-in with the same interface; the class under test is
tcnn_b200.dp.DataParallelTrainer, which bench.py --gpus N uses with the real trainer over NCCL.)"""
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = json.load(open(os.path.join(ROOT, "tests", "golden", "configs", "hash3d_small.json")))
B_GLOBAL = 1024
STEPS = 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "tiny-cuda-nn_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from tcnn_b200.dp import DataParallelTrainer

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = ob.default_rng(1337)
        x = ob.generate_random_uniform(rng, B_GLOBAL * 3).reshape(B_GLOBAL, 3)
        y = ob.make_targets(x, 3)
        dp = DataParallelTrainer(ob.OracleShardTrainer(3, 3, CFG))
        assert (dp.world, dp.rank) == (world, rank)
        lo, hi = dp.shard(B_GLOBAL)
        losses = []
        for _ in range(STEPS):
            dp.training_step(torch.from_numpy(x[lo:hi]), torch.from_numpy(y[lo:hi]))
            losses.append(dp.loss())
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), params=dp.trainer.m.params_fp32, steps=dp.trainer.m.steps, losses=np.array(losses), shard=np.array([lo, hi]))
        with pytest.raises(ValueError):
            dp.shard(B_GLOBAL + 256)  # not divisible into 256-multiples per rank
    finally:
        dist.destroy_process_group()


def test_two_rank_data_parallel_matches_single_process(tmp_path):
    import oracle_binding as ob

    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    # shards partition the batch
    assert list(r0["shard"]) == [0, 512] and list(r1["shard"]) == [512, 1024]
    # replicas stay bit-identical (same reduced gradients -> same Adam step, including the zero-gradient skip)
    assert np.array_equal(r0["params"].view(np.uint32), r1["params"].view(np.uint32))
    assert np.array_equal(r0["steps"], r1["steps"])
    assert np.allclose(r0["losses"], r1["losses"], rtol=0, atol=0)

    # single process, full batch
    rng = ob.default_rng(1337)
    x = ob.generate_random_uniform(rng, B_GLOBAL * 3).reshape(B_GLOBAL, 3)
    y = ob.make_targets(x, 3)
    ref = ob.OracleModel(3, 3, CFG)
    ref_losses = [ref.training_step(x, y) for _ in range(STEPS)]
    assert np.allclose(r0["losses"], ref_losses, rtol=1e-6)
    # gradients are exact sums in both cases -> parameters agree to fp32 rounding of one Adam step
    assert np.array_equal(r0["steps"], ref.steps)
    assert np.abs(r0["params"] - ref.params_fp32).max() < 1e-6
