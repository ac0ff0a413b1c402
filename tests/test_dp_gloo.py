"""Data-parallel host logic on CPU: two (and four) gloo ranks, each holding a replica, must reproduce the single-process step.
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


def _worker(rank, world, port, out_dir, shard_optimizer):
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
        dp = DataParallelTrainer(ob.OracleShardTrainer(3, 3, CFG), shard_optimizer=shard_optimizer)
        assert (dp.world, dp.rank, dp.shard_optimizer) == (world, rank, shard_optimizer)
        lo, hi = dp.shard(B_GLOBAL)
        losses = []
        for _ in range(STEPS):
            dp.training_step(torch.from_numpy(x[lo:hi]), torch.from_numpy(y[lo:hi]))
            losses.append(dp.loss())
        stale = dp.trainer.m.params_fp32.copy()  # sharded optimizer: masters of the other ranks' slices are not current yet
        dp.sync_full_precision()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), params=dp.trainer.m.params_fp32, params_before_sync=stale, params16=dp.trainer.m.params_fp16,
                 steps=dp.trainer.m.steps, losses=np.array(losses), shard=np.array([lo, hi]), owned=np.array(dp.owned_range()))
        with pytest.raises(ValueError):
            dp.shard(B_GLOBAL + 256)  # not divisible into 256-multiples per rank
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shard_optimizer,world", [(False, 2), (True, 2), (True, 4)])
def test_data_parallel_matches_single_process(tmp_path, shard_optimizer, world):
    import oracle_binding as ob

    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), shard_optimizer), nprocs=world, join=True)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    r0, r1 = ranks[0], ranks[1]
    # shards partition the batch
    per = B_GLOBAL // world
    assert [list(r["shard"]) for r in ranks] == [[i * per, (i + 1) * per] for i in range(world)]
    # replicas stay bit-identical (same reduced gradients -> same Adam step, including the zero-gradient skip): the working
    # (fp16) parameters after every step, the fp32 masters once the slices have been exchanged
    for r in ranks[1:]:
        assert np.array_equal(r0["params16"], r["params16"])
        assert np.array_equal(r0["params"].view(np.uint32), r["params"].view(np.uint32))
        assert np.allclose(r0["losses"], r["losses"], rtol=0, atol=0)

    # single process, full batch
    rng = ob.default_rng(1337)
    x = ob.generate_random_uniform(rng, B_GLOBAL * 3).reshape(B_GLOBAL, 3)
    y = ob.make_targets(x, 3)
    ref = ob.OracleModel(3, 3, CFG)
    ref_losses = [ref.training_step(x, y) for _ in range(STEPS)]
    assert np.allclose(r0["losses"], ref_losses, rtol=1e-6)
    # gradients are exact sums in both cases -> parameters agree to fp32 rounding of one Adam step
    assert np.abs(r0["params"] - ref.params_fp32).max() < 1e-6
    assert np.array_equal(r0["params16"], ref.params_fp16)
    n_mlp, n = ref.n_mlp, ref.n_params
    if not shard_optimizer:
        assert np.array_equal(r0["steps"], r1["steps"]) and np.array_equal(r0["steps"], ref.steps)
        assert list(r0["owned"]) == [0, n]
    else:
        # the padded parameter vector is cut into two equal slices; rank 0's holds the network weights. The owner's
        # per-parameter step counters (adam.h:100) match the single-process run.
        chunk = ((n + 511) // 512 * 512) // world
        assert chunk >= n_mlp
        assert [list(r["owned"]) for r in ranks] == [[i * chunk, max(0, min(chunk, n - i * chunk))] for i in range(world)]
        for r in ranks:
            b, c = r["owned"]
            assert np.array_equal(r["steps"][b : b + c], ref.steps[b : b + c])
        # before the exchange a rank's masters of the OTHER ranks' slices are stale (still the initial values there)
        other = slice(chunk, 2 * chunk)
        assert not np.array_equal(r0["params_before_sync"][other], r0["params"][other])
