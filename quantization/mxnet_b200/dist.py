"""Data-parallel plumbing for calibration and QAT: one process per GPU, batch sharded by sample,
weights replicated.  Every statistic on this path is a max, an integer sum or a mean of
per-sample maxima, so the collectives below make an N-GPU run reproduce the single-GPU result:

  first-batch maxima (KL)      all_reduce(MAX)   [L]          exact
  per-batch histogram counts   all_reduce(SUM)   [S, L, bins+1]  exact (32-bit on the wire when the totals fit, else
                                                              int64); S (8) batches per collective, asynchronous
  per-sample input maxima      all_gather        [N]          exact; the Kahan mean then runs on every rank
  QAT gradients                all_reduce(SUM)/R one flat fp32 bucket

All messages are tiny (<= 442 KB) except the gradient bucket; they go through torch.distributed
(NCCL over NVLink on GPUs, gloo in the CPU tests).  Nothing here launches a kernel of its own.
"""
import torch
import torch.distributed as dist

__all__ = ["active_group", "shard_batch", "sync_first_batch_minmax", "sync_counts", "CountsRing", "gather_per_sample",
           "GradBucket", "broadcast_parameters", "enable_data_parallel", "disable_data_parallel"]


def active_group(group=None):
    """The process group to use, or None when running single-process."""
    if group is not None:
        return group
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.group.WORLD
    return None


def shard_batch(x, rank=None, world=None, group=None):
    """Contiguous N/R samples for this rank (N must divide evenly, as per-GPU batches are fixed)."""
    g = active_group(group)
    if rank is None:
        rank = dist.get_rank(g) if g is not None else 0
    if world is None:
        world = dist.get_world_size(g) if g is not None else 1
    n = x.shape[0]
    if n % world != 0:
        raise ValueError("batch of %d samples does not split over %d ranks" % (n, world))
    per = n // world
    return x[rank * per:(rank + 1) * per]


def sync_first_batch_minmax(minmax, group=None):
    """minmax: [L, 2] = {min, max} per layer of this rank's shard of batch 0 -> global, in place."""
    g = active_group(group)
    if g is None:
        return minmax
    mn = minmax[:, 0].contiguous()
    mx = minmax[:, 1].contiguous()
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=g)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=g)
    minmax[:, 0].copy_(mn)
    minmax[:, 1].copy_(mx)
    return minmax


def sync_counts(counts, group=None):
    """counts: int64 [L, bins+1] of this rank's shard -> global integer counts, in place (one collective)."""
    g = active_group(group)
    if g is not None:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=g)
    return counts


class CountsRing:
    """Per-batch integer counts of up to ``slots`` batches, all-reduced TOGETHER and off the critical path.

    The reference adds ``float32(counts)`` of one batch after the other (distribution_calibrate.py:47,103-104),
    so the counts of a global batch must be summed over the ranks before they are folded -- but nothing needs
    them before the KL search.  Keeping every batch's counts in its own slot and replaying the float32 adds in
    batch order after ONE collective per ``slots`` batches gives the same bits as an all-reduce per batch, without
    a latency-bound collective in every step (and one fold launch per ``slots`` batches instead of one per batch).

    Two rings take turns: while the histogram kernels fill one, the other one's sum-all-reduce runs asynchronously on
    the communicator's stream; its fold is queued when the next ring is full (or at ``flush()``), so no histogram
    launch ever waits for a collective.  On the wire the counts are 32-bit whenever the per-bin totals of a global
    batch provably fit (``max_count`` < 2^32 summed over the ranks): half the message.  Both are invisible in the
    result (integer sums are exact; the folds still run in batch order).

    ``accumulate(counts_2d [S, n], first)`` is the fold (ops.hist_accumulate on the GPU); ``on_reduced(counts_3d)``
    sees the global counts before they are folded (deferred 2049th-bin check).
    """

    def __init__(self, n_layers, n_bins, device, accumulate, group=None, slots=8, on_reduced=None, max_count=None,
                 local=False):
        self.group = None if local else active_group(group)      # local: no exchange even with ranks present
        self.slots = max(1, int(slots))
        world = dist.get_world_size(self.group) if self.group is not None else 1
        # max_count: an upper bound on one rank's count in any bin of any batch (= its largest layer input)
        narrow = max_count is not None and int(max_count) * world < (1 << 32) and torch.device(device).type == "cuda"
        self.dtype = torch.int32 if narrow else torch.int64     # int32 carries uint32 sums (two's complement add)
        self.rings = [torch.zeros(self.slots, n_layers, n_bins, dtype=self.dtype, device=device)
                      for _ in range(2 if self.group is not None else 1)]
        self.cur = 0
        self.used = 0
        self.flushed = 0
        self.accumulate = accumulate
        self.on_reduced = on_reduced
        self._inflight = None          # (work handle, ring index, slots used)

    @property
    def ring(self):
        return self.rings[self.cur]

    def prime(self, slot_counts=None):
        """Run the collective once on the (all-zero) rings for every message size that will occur.  NCCL connects an
        algorithm/protocol the first time a message size selects it (runtime connect): on 8 GPUs the first 14 MB
        all-reduce of a calibration took 11 ms instead of ~0.1 ms.  Zeros stay zeros, so this changes nothing."""
        if self.group is None:
            return
        assert self.used == 0 and self._inflight is None, "prime() must not see collected counts"
        for n in sorted({min(max(int(c), 1), self.slots) for c in (slot_counts or [self.slots])}):
            for ring in self.rings:
                dist.all_reduce(ring[:n], op=dist.ReduceOp.SUM, group=self.group)

    def slot(self):
        """[n_layers, n_bins] zeroed counters for the batch being collected."""
        return self.ring[self.used]

    def commit(self):
        """The current slot holds a complete batch."""
        self.used += 1
        if self.used == self.slots:
            self._rotate()

    def _retire(self):
        """Fold the ring whose all-reduce was started earlier (the current stream waits for it; the host does not)."""
        if self._inflight is None:
            return
        work, idx, used = self._inflight
        self._inflight = None
        if work is not None:
            work.wait()
        part = self.rings[idx][:used]
        if self.on_reduced is not None:
            self.on_reduced(part)
        self.accumulate(part.view(used, -1), self.flushed == 0)     # zeroes the slots again
        self.flushed += used

    def _rotate(self):
        if self.used == 0:
            return
        self._retire()                  # batch order: the older ring is folded first (and is free again)
        part = self.ring[:self.used]
        work = None
        if self.group is not None:
            work = dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._inflight = (work, self.cur, self.used)
        if self.group is None:
            self._retire()              # single process: nothing to overlap, fold right away
        else:
            self.cur = 1 - self.cur
        self.used = 0

    def flush(self):
        """Everything collected so far is folded into the histograms when this returns (stream order)."""
        self._rotate()
        self._retire()


def gather_per_sample(per_sample, group=None):
    """[N/R] per-sample maxima of this rank -> [N] in rank (= sample) order on every rank."""
    g = active_group(group)
    if g is None:
        return per_sample
    world = dist.get_world_size(g)
    out = torch.empty(world * per_sample.numel(), dtype=per_sample.dtype, device=per_sample.device)
    dist.all_gather_into_tensor(out, per_sample.contiguous(), group=g)
    return out


def broadcast_parameters(net, src=0, group=None):
    g = active_group(group)
    if g is None:
        return
    with torch.no_grad():
        for t in list(net.parameters()) + list(net.buffers()):
            dist.broadcast(t.data, src=src, group=g)


class GradBucket:
    """All gradients of a net live in ONE flat fp32 buffer (each ``p.grad`` is a view into it, the way DDP's
    gradient_as_bucket_view works), so a QAT step needs a single all-reduce and no packing copies.
    Use ``optimizer.zero_grad(set_to_none=False)`` so that the views survive.

    ``net=`` (a converted net under :func:`enable_data_parallel`): the per-sample input maxima of the step ride in
    the tail of the same buffer -- every rank fills its own [L, N/R] slice of an otherwise zero [R, L, N/R] block,
    so the sum-all-reduce IS the all-gather (x + 0 is exact) -- and ``net.update_ema()`` calls made earlier in the
    step are completed right after the collective.  The step then has one rendezvous between the ranks instead
    of two (the second one, in the middle of the step, costs about 0.5 ms of rank skew per step on 2-4 GPUs)."""

    def __init__(self, params, group=None, net=None, align=32):
        self.all_params = [p for p in params if p.requires_grad]
        self.params = list(self.all_params)
        self.group = group
        # every gradient view starts on a 128-byte boundary of the flat buffer (vectorised access for every kernel
        # that touches a gradient)
        self.align = max(1, int(align))
        self._layout()
        self.flat = None
        self.net = net
        self._deferred_ema = []
        if net is not None:
            net._fq_grad_bucket = self

    def _layout(self):
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off = (off + p.numel() + self.align - 1) // self.align * self.align
        self.numel = off

    def attach(self, tail_numel=0):
        """Make every gradient a view of the flat buffer.  Called after a backward pass (all_reduce_mean() does it
        on first use), parameters that received no gradient stay out of the bucket: a bypassed BatchNorm still owns
        trainable weight / bias that nothing reads, and giving them zero gradients would make the optimizer step
        through them on every iteration (measured on MobileNetV2 with fake-BN: 104 of 317 parameters, +0.5 ms per
        7.8 ms step)."""
        if any(p.grad is not None for p in self.all_params):
            used = [p for p in self.all_params if p.grad is not None]
            if len(used) != len(self.params) or any(a is not b for a, b in zip(used, self.params)):
                self.params = used
                self._layout()
        dev = self.params[0].device
        self.flat = torch.zeros(self.numel + tail_numel, dtype=torch.float32, device=dev)
        for p, off in zip(self.params, self.offsets):
            view = self.flat[off:off + p.numel()].view(p.shape)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
        return self

    def _attached(self):
        if self.flat is None or not self.params:
            return False
        p = self.params[0]
        return p.grad is not None and p.grad.data_ptr() == self.flat.data_ptr()

    # ---- input ranges riding along -----------------------------------------------------------------------------
    def _pending(self):
        if self.net is None:
            return []
        from .quantize.convert.convert_conv2d import pending_range_blocks
        return pending_range_blocks(self.net.collect_quantized_blocks())

    def takes_over_ema(self):
        """True when update_ema() should wait for all_reduce_mean(): training step, ranges still shard-local."""
        return (self.net is not None and torch.is_grad_enabled() and active_group(self.group) is not None
                and bool(self._pending()))

    def defer_ema(self, momentum):
        self._deferred_ema.append(momentum)

    def all_reduce_mean(self):
        g = active_group(self.group)
        if g is None or not self.params:
            return
        world, rank = dist.get_world_size(g), dist.get_rank(g)
        todo = self._pending()
        n = todo[0]._fq_per_sample.numel() if todo else 0
        tail_numel = world * len(todo) * n
        if not self._attached() or self.flat.numel() != self.numel + tail_numel:
            self.attach(tail_numel)
        tail = self.flat[self.numel:].view(world, len(todo), n) if todo else None
        if todo:
            from .quantize.convert.convert_conv2d import local_range_arena
            local_range_arena(todo, into=tail[rank])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=g)
        self.flat[:self.numel].div_(world)
        if todo:
            from .quantize.convert.convert_conv2d import finish_global_ranges
            finish_global_ranges(todo, tail)
            tail.zero_()          # the other ranks' slices must be zero again; ours is rewritten by the next forward
        deferred, self._deferred_ema = self._deferred_ema, []
        for momentum in deferred:
            self.net.update_ema(momentum)


def enable_data_parallel(net, group=None):
    """Make every converted block compute its ONLINE input range over the global batch: the block keeps
    its shard's per-sample maxima, all-gathers them and runs the reference's Kahan mean on every rank."""
    g = active_group(group)
    for m in net.collect_quantized_blocks():
        m._fq_dist_group = g
    return net


def disable_data_parallel(net):
    for m in net.collect_quantized_blocks():
        m._fq_dist_group = None
    return net
