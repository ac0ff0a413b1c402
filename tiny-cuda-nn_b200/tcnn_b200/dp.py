"""Sample-sharded data parallelism for the training step -- host-side logic, backend agnostic.

The reference is single-GPU; this is the new multi-GPU path. One process per GPU (torchrun), every rank holds a full
replica of the working-precision parameters. A step is:

    1. each rank runs fwd+loss+bwd on ITS shard with the loss normalised over the GLOBAL batch
       (`training_step_shard`), so per-rank gradients are partial sums of the single-GPU gradient;
    2. the gradients are summed over the ranks and Adam is applied -- in one of two ways:

       shard_optimizer=True (default, ZeRO-1 style). The parameter vector [network weights | grid table], padded to a multiple
       of 512, is cut into `world` equal slices; rank r owns slice r:
         reduce-scatter(sum) of the fp16 gradient vector -> rank r holds the reduced gradients of its slice;
         Adam on that slice only -- the optimizer pass, the largest HBM consumer of the step (36 B/parameter), shrinks by the
         world size;
         all-gather of the updated working-precision (fp16) slices, launched on a side stream: the next step's binning pass
         (which reads only positions) overlaps it, only the fused kernel waits for it.
         Two collectives per step, the same bytes on the wire as one all-reduce. fp32 master parameters and Adam moments of a
         slice live on its owner only; `sync_full_precision()` all-gathers the masters (before serialising, or to read them).

       shard_optimizer=False: all-reduce of both gradient buffers (fp16 table, fp32 network accumulator), then the same full
       Adam step on every replica.

    Either way the zero-gradient skip of adam.h:79-82 is evaluated on the REDUCED gradients, which keeps the replicas'
    working parameters bit-identical.

`trainer` is anything with the methods used below (the CUDA trainer in production; the CPU tests drive the same class with an
oracle-backed stand-in over gloo):
    training_step_shard(x, y, global_batch, run_optimizer), optimizer_step(ranges=None), loss(), gradient_buffers(),
    and for the sharded optimizer: sharded_buffers() -> {"grads", "params", "masters": padded whole-vector tensors,
    "n_params", "n_matrix"}, finalize_gradients(), and optionally wait_before_compute(event) (CUDA only).
"""
import torch.distributed as dist

# slices handed to Adam / the collectives start on multiples of this many parameters (vectorised kernels, 16-byte alignment)
SLICE_GRANULARITY = 8


class DataParallelTrainer:
    def __init__(self, trainer, group=None, shard_optimizer=True, native=True, peer_memory=True):
        self.trainer = trainer
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.shard_optimizer = bool(shard_optimizer) and self.world > 1 and hasattr(trainer, "sharded_buffers")
        if self.shard_optimizer:
            b = trainer.sharded_buffers()
            n_pad = b["grads"].numel()
            if n_pad % (SLICE_GRANULARITY * self.world) != 0 or n_pad // self.world < b["n_matrix"]:
                self.shard_optimizer = False  # slices must be aligned, and the network weights must sit inside slice 0
        self._masters_synced = True
        self._comm_stream = None
        # Production engine: the whole step (collectives included) is one call into libtcnn_b200, which talks to NCCL itself --
        # a Python-driven step costs ~0.4 ms of host time, more than the step takes on the device. The torch.distributed
        # version below is the same logic; it serves backends without NCCL (the gloo CPU tests) and as readable reference.
        self.native = False
        self.engine = "single" if self.world == 1 else "torch.distributed"
        self._symm = None
        if self.world > 1 and native and hasattr(trainer, "dp_native_init") and dist.get_backend(group) == "nccl":
            self.shard_optimizer = trainer.dp_native_init(group, shard_optimizer)
            self.native = True
            self.engine = "nccl"
            # Peer-memory engine: the working parameters and the gradient vector move into a symmetric allocation every rank has
            # mapped (torch's symmetric memory does the rendezvous: peer pointers + the NVLS multicast mapping); the library then
            # replaces reduce-scatter -> Adam -> all-gather by ONE kernel between two NVLink flag barriers (misc_kernels.cu).
            if peer_memory and self.shard_optimizer and self.world <= 8 and hasattr(trainer, "dp_attach_symmetric"):
                try:
                    self.engine = self._attach_symmetric(group)
                except Exception as e:  # noqa: BLE001 -- no symmetric memory on this system: stay on the NCCL engine, and say so
                    import warnings

                    warnings.warn(f"tcnn_b200: peer-memory data-parallel engine unavailable ({e!r}); using NCCL collectives")
                ok = [self.engine]
                dist.all_gather_object(all_engines := [None] * self.world, ok[0], group=group)
                if len(set(all_engines)) != 1:
                    raise RuntimeError(f"data-parallel ranks disagree on the engine: {all_engines}")

    def _attach_symmetric(self, group):
        import torch
        import torch.distributed._symmetric_memory as symm_mem

        n_bytes = self.trainer.dp_symmetric_bytes()
        dev = torch.device("cuda", torch.cuda.current_device())
        buf = symm_mem.empty(n_bytes, dtype=torch.uint8, device=dev)
        hdl = symm_mem.rendezvous(buf, group=group if group is not None else dist.group.WORLD)
        peers = [int(p) for p in hdl.buffer_ptrs]
        mc = int(hdl.multicast_ptr) if getattr(hdl, "multicast_ptr", 0) else 0
        assert peers[self.rank] == buf.data_ptr()
        self.trainer.dp_attach_symmetric(peers, mc, n_bytes)
        self._symm = (buf, hdl)  # keeps the mapping alive
        torch.cuda.synchronize()
        dist.barrier(group=group)
        return "peer-memory-multicast" if mc else "peer-memory-p2p"

    def shard(self, n_global):
        """[begin, end) of this rank's contiguous shard of a global batch; shards must stay multiples of 256."""
        if n_global % (self.world * 256) != 0:
            raise ValueError(f"global batch {n_global} must be a multiple of 256 * world_size ({self.world})")
        per = n_global // self.world
        return self.rank * per, (self.rank + 1) * per

    # ------------------------------------------------------------------ parameter slices of the sharded optimizer
    def owned_range(self):
        """(begin, count) of the parameters this rank's optimizer state is authoritative for."""
        if not self.shard_optimizer:
            g = self.trainer.gradient_buffers()
            return 0, sum(int(t.numel()) for t in g) if len(g) > 1 else int(g[0].numel())
        b = self.trainer.sharded_buffers()
        chunk = b["grads"].numel() // self.world
        begin = self.rank * chunk
        return begin, max(0, min(chunk, b["n_params"] - begin))

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

    def _all_gather_overlapped(self, whole, own):
        """All-gather on a side stream; the trainer's next fused kernel waits for it, the binning pass in front of it does not."""
        import torch

        if not (whole.is_cuda and hasattr(self.trainer, "wait_before_compute")):
            self._all_gather(whole, own)
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream()
        self._comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._comm_stream):
            self._all_gather(whole, own)
            done = self._comm_stream.record_event()
        self.trainer.wait_before_compute(done)

    # ------------------------------------------------------------------ the step
    def training_step(self, x_shard, y_shard):
        """x_shard / y_shard: this rank's samples. Returns nothing; the global loss is `loss()`."""
        t = self.trainer
        global_batch = x_shard.shape[0] * self.world
        if self.world == 1:
            t.training_step_shard(x_shard, y_shard, global_batch, run_optimizer=True)
            return
        if self.native:
            t.dp_training_step(x_shard, y_shard, global_batch)
            self._masters_synced = not self.shard_optimizer
            return
        t.training_step_shard(x_shard, y_shard, global_batch, run_optimizer=False)
        if not self.shard_optimizer:
            for buf in t.gradient_buffers():
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            t.optimizer_step()
            return

        t.finalize_gradients()
        b = t.sharded_buffers()
        chunk = b["grads"].numel() // self.world
        lo = self.rank * chunk
        self._reduce_scatter(b["grads"], b["grads"][lo : lo + chunk])
        begin, count = self.owned_range()
        # an empty slice (count == 0: this rank owns only padding) still counts as an optimizer step, as in the native
        # path (model.cu dp_training_step): AdaBound's learning-rate bounds depend on the step counter
        t.optimizer_step(ranges=[(begin, count) if count else (0, 0)])
        self._all_gather_overlapped(b["params"], b["params"][lo : lo + chunk])
        self._masters_synced = False

    def sync_full_precision(self):
        """All-gather the fp32 master parameters (the owner of a slice holds its current values)."""
        if not self.shard_optimizer or self._masters_synced:
            return
        if self.native:
            self.trainer.dp_sync_full_precision()
            self._masters_synced = True
            return
        m = self.trainer.sharded_buffers()["masters"]
        chunk = m.numel() // self.world
        self._all_gather(m, m[self.rank * chunk : (self.rank + 1) * chunk])
        self._masters_synced = True

    def close(self):
        """Tear the native communicators down (before torch.distributed's process group is destroyed)."""
        if self.native:
            self.trainer.dp_finish()
            self.native = False
            self._symm = None

    def loss(self):
        """Sum of the ranks' partial losses == the single-GPU loss of the global batch."""
        import torch

        v = torch.tensor([self.trainer.loss()], dtype=torch.float64)
        if self.world > 1:
            dev = self.trainer.device() if hasattr(self.trainer, "device") else "cpu"
            v = v.to(dev)
            dist.all_reduce(v, group=self.group)
        return float(v.item())
