"""Sample-sharded data parallelism for the training step -- host-side logic, backend agnostic.

The reference is single-GPU; this is the new multi-GPU path. One process per GPU (torchrun), every rank holds a full
replica. A step is:

    1. each rank runs fwd+loss+bwd on ITS shard with the loss normalised over the GLOBAL batch
       (`training_step_shard`), so per-rank gradients are partial sums of the single-GPU gradient;
    2. one all-reduce(sum) over the grid-gradient table (fp16) and one over the MLP weight-gradient accumulator (fp32);
    3. every rank applies the same Adam step to its replica (`optimizer_step`) -- the zero-gradient skip of adam.h:79-82 is
       evaluated on the REDUCED gradients, which keeps the replicas bit-identical.

`trainer` is anything with training_step_shard / optimizer_step / gradient_buffers (the CUDA trainer in production; the
CPU tests drive the same class with an oracle-backed stand-in over gloo).
"""
import torch.distributed as dist


class DataParallelTrainer:
    def __init__(self, trainer, group=None):
        self.trainer = trainer
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def shard(self, n_global):
        """[begin, end) of this rank's contiguous shard of a global batch; shards must stay multiples of 256."""
        if n_global % (self.world * 256) != 0:
            raise ValueError(f"global batch {n_global} must be a multiple of 256 * world_size ({self.world})")
        per = n_global // self.world
        return self.rank * per, (self.rank + 1) * per

    def training_step(self, x_shard, y_shard):
        """x_shard / y_shard: this rank's samples. Returns nothing; the global loss is `loss()`."""
        global_batch = x_shard.shape[0] * self.world
        if self.world == 1:
            self.trainer.training_step_shard(x_shard, y_shard, global_batch, run_optimizer=True)
            return
        self.trainer.training_step_shard(x_shard, y_shard, global_batch, run_optimizer=False)
        for buf in self.trainer.gradient_buffers():
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        self.trainer.optimizer_step()

    def loss(self):
        """Sum of the ranks' partial losses == the single-GPU loss of the global batch."""
        import torch

        v = torch.tensor([self.trainer.loss()], dtype=torch.float64)
        if self.world > 1:
            dev = self.trainer.device() if hasattr(self.trainer, "device") else "cpu"
            v = v.to(dev)
            dist.all_reduce(v, group=self.group)
        return float(v.item())
