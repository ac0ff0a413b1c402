"""Sample-sharded data parallelism for the training step -- host-side logic, backend agnostic.

The reference is single-GPU; this is the new multi-GPU path. One process per GPU (torchrun), every rank holds a full
replica of the working-precision parameters. A step is:

    1. each rank runs fwd+loss+bwd on ITS shard with the loss normalised over the GLOBAL batch
       (`training_step_shard`), so per-rank gradients are partial sums of the single-GPU gradient;
    2. the gradients are summed over the ranks and Adam is applied -- in one of two ways:

       shard_optimizer=True (default, ZeRO-1 style):
         reduce-scatter(sum) of the grid-gradient table: rank r receives the reduced gradients of ITS slice of the table;
         all-reduce of the (tiny) network weight gradients and of the few table entries left over by the equal split;
         Adam on the network weights (every rank, identical) and on the rank's own table slice only -- the optimizer pass, the
         largest HBM consumer of the step (36 B/parameter), shrinks by the world size;
         all-gather of the updated working-precision (fp16) table slices.
         Same bytes on the wire as an all-reduce. fp32 master parameters and Adam moments of a slice live on its owner only;
         `sync_full_precision()` all-gathers the masters (before serialising, or to read them anywhere).

       shard_optimizer=False: all-reduce of both gradient buffers, then the same full Adam step on every replica.

    Either way the zero-gradient skip of adam.h:79-82 is evaluated on the REDUCED gradients, which keeps the replicas'
    working parameters bit-identical.

`trainer` is anything with the methods used below (the CUDA trainer in production; the CPU tests drive the same class with an
oracle-backed stand-in over gloo):
    training_step_shard(x, y, global_batch, run_optimizer), optimizer_step(ranges=None), loss(), gradient_buffers(),
    and for the sharded optimizer: shardable_gradients() -> (tensor, first_param), replicated_gradients() -> [tensor],
    params() -> working-precision tensor of all parameters, params_full_precision() -> fp32 tensor of all parameters.
"""
import torch.distributed as dist

# slices handed to Adam / the collectives start on multiples of this many parameters (vectorised kernels, 16-byte alignment)
SLICE_GRANULARITY = 8


class DataParallelTrainer:
    def __init__(self, trainer, group=None, shard_optimizer=True):
        self.trainer = trainer
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.shard_optimizer = bool(shard_optimizer) and self.world > 1 and hasattr(trainer, "shardable_gradients")
        self._masters_synced = True

    def shard(self, n_global):
        """[begin, end) of this rank's contiguous shard of a global batch; shards must stay multiples of 256."""
        if n_global % (self.world * 256) != 0:
            raise ValueError(f"global batch {n_global} must be a multiple of 256 * world_size ({self.world})")
        per = n_global // self.world
        return self.rank * per, (self.rank + 1) * per

    # ------------------------------------------------------------------ parameter slices of the sharded optimizer
    def slice_layout(self, n_shardable):
        """Equal split of `n_shardable` table parameters: (slice length, number of left-over parameters at the end)."""
        chunk = (n_shardable // (SLICE_GRANULARITY * self.world)) * SLICE_GRANULARITY
        return chunk, n_shardable - chunk * self.world

    def owned_ranges(self):
        """Parameter ranges [(begin, count)] this rank's optimizer state is authoritative for."""
        g, first = self.trainer.shardable_gradients()
        if not self.shard_optimizer:
            return [(0, first + g.numel())]
        chunk, tail = self.slice_layout(g.numel())
        ranges = [(0, first), (first + self.rank * chunk, chunk)]
        if tail:
            ranges.append((first + chunk * self.world, tail))
        return [r for r in ranges if r[1] > 0]

    # ------------------------------------------------------------------ collectives (NCCL; gloo fallbacks for the CPU tests)
    def _reduce_scatter(self, whole, own):
        if dist.get_backend(self.group) == "gloo":  # gloo has no reduce-scatter: reduce everything, keep the own slice
            dist.all_reduce(whole, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.reduce_scatter_tensor(own, whole, op=dist.ReduceOp.SUM, group=self.group)  # in place: own is whole[rank]

    def _all_gather(self, whole, own):
        if dist.get_backend(self.group) == "gloo":
            import torch

            parts = list(whole.view(torch.uint8).chunk(self.world))  # byte view: gloo has no 16-bit types
            dist.all_gather(parts, own.view(torch.uint8).clone(), group=self.group)
        else:
            dist.all_gather_into_tensor(whole, own, group=self.group)  # in place: own is whole[rank]

    # ------------------------------------------------------------------ the step
    def training_step(self, x_shard, y_shard):
        """x_shard / y_shard: this rank's samples. Returns nothing; the global loss is `loss()`."""
        t = self.trainer
        global_batch = x_shard.shape[0] * self.world
        if self.world == 1:
            t.training_step_shard(x_shard, y_shard, global_batch, run_optimizer=True)
            return
        t.training_step_shard(x_shard, y_shard, global_batch, run_optimizer=False)
        if not self.shard_optimizer:
            for buf in t.gradient_buffers():
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            t.optimizer_step()
            return

        grads, first = t.shardable_gradients()
        chunk, tail = self.slice_layout(grads.numel())
        lo = self.rank * chunk
        if chunk:
            self._reduce_scatter(grads[: chunk * self.world], grads[lo : lo + chunk])
        for buf in t.replicated_gradients():
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        if tail:
            dist.all_reduce(grads[chunk * self.world :], op=dist.ReduceOp.SUM, group=self.group)
        t.optimizer_step(ranges=self.owned_ranges())
        if chunk:
            table = t.params()[first : first + chunk * self.world]
            self._all_gather(table, table[lo : lo + chunk])
        self._masters_synced = False

    def sync_full_precision(self):
        """All-gather the fp32 master parameters of the table slices (their owners hold the current values)."""
        if not self.shard_optimizer or self._masters_synced:
            return
        grads, first = self.trainer.shardable_gradients()
        chunk, _ = self.slice_layout(grads.numel())
        if chunk:
            masters = self.trainer.params_full_precision()[first : first + chunk * self.world]
            self._all_gather(masters, masters[self.rank * chunk : (self.rank + 1) * chunk])
        self._masters_synced = True

    def loss(self):
        """Sum of the ranks' partial losses == the single-GPU loss of the global batch."""
        import torch

        v = torch.tensor([self.trainer.loss()], dtype=torch.float64)
        if self.world > 1:
            dev = self.trainer.device() if hasattr(self.trainer, "device") else "cpu"
            v = v.to(dev)
            dist.all_reduce(v, group=self.group)
        return float(v.item())
