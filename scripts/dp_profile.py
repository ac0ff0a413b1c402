"""Phase timing of the data-parallel step (run under torchrun, one rank per GPU): device time of each phase with a
synchronize in between (so nothing overlaps), then the real pipelined step for comparison."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-cuda-nn_b200"))
import torch
import torch.distributed as dist

import tcnn_b200
from tcnn_b200.dp import DataParallelTrainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = 1 << 18
cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "configs", "headline.json")))
model = tcnn_b200.create_from_config(3, 3, cfg)
t = model.trainer
dp = DataParallelTrainer(t, native=False)  # phase timing drives the torch.distributed engine
g = torch.Generator(device="cuda").manual_seed(1 + rank)
x, y = torch.rand(B, 3, device="cuda", generator=g), torch.rand(B, 3, device="cuda", generator=g)
for _ in range(10):
    dp.training_step(x, y)
torch.cuda.synchronize()
dist.barrier()


def timed(fn, n=20):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    tot, host = 0.0, 0.0
    for _ in range(n):
        torch.cuda.synchronize()
        dist.barrier()
        h0 = time.perf_counter()
        ev[0].record()
        fn()
        ev[1].record()
        host += time.perf_counter() - h0
        torch.cuda.synchronize()
        tot += ev[0].elapsed_time(ev[1])
    return tot / n, host / n * 1e3


b = t.sharded_buffers()
chunk = b["grads"].numel() // world
lo = rank * chunk
res = {}
res["shard_step"] = timed(lambda: t.training_step_shard(x, y, B * world, run_optimizer=False))
res["finalize"] = timed(lambda: t.finalize_gradients())
res["reduce_scatter"] = timed(lambda: dp._reduce_scatter(b["grads"], b["grads"][lo : lo + chunk]))
res["all_reduce_same_bytes"] = timed(lambda: dist.all_reduce(b["grads"]))
begin, count = dp.owned_range()
t.training_step_shard(x, y, B * world, run_optimizer=False)
res["adam_slice"] = timed(lambda: t.optimizer_step(ranges=[(begin, count)]))
res["all_gather"] = timed(lambda: dp._all_gather(b["params"], b["params"][lo : lo + chunk]))
res["all_gather_side_stream"] = timed(lambda: dp._all_gather_overlapped(b["params"], b["params"][lo : lo + chunk]))
res["full_step"] = timed(lambda: dp.training_step(x, y))
# pipelined: many steps back to back
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
h0 = time.perf_counter()
e0.record()
for _ in range(100):
    dp.training_step(x, y)
e1.record()
host_ms = (time.perf_counter() - h0) * 10
torch.cuda.synchronize()
res["pipelined_step"] = (e0.elapsed_time(e1) / 100, host_ms)
dpn = DataParallelTrainer(t, native=True)
for _ in range(10):
    dpn.training_step(x, y)
torch.cuda.synchronize()
dist.barrier()
h0 = time.perf_counter()
e0.record()
for _ in range(100):
    dpn.training_step(x, y)
e1.record()
host_ms = (time.perf_counter() - h0) * 10
torch.cuda.synchronize()
res["pipelined_step_native"] = (e0.elapsed_time(e1) / 100, host_ms)
dpn.close()
if rank == 0:
    print(json.dumps({"dp_profile_ms(device, host)": {k: [round(v[0], 4), round(v[1], 4)] for k, v in res.items()}, "world": world}))
dist.destroy_process_group()
