"""Conv2D converter (convert_conv2d.py of the reference) over torch.nn.Conv2d.

The patched forward keeps the reference's control flow and per-block state (``quantize_args``,
``fixed_params`` tri-state, ``enable_quantize``, ``quantize_input``, ``quantize_input_offline``,
``current_input_max``, Parameter ``input_max``; fake-BN ``gamma/beta/running_mean/running_var``),
but each tensor goes through one or two fused kernel launches and nothing is read back to the host:
``current_input_max`` is a (1,) device tensor instead of a Python float.
"""
import types
from collections import namedtuple

import torch
from torch import nn

from ... import ops
from .wino_matrix import winograd_matrices

__all__ = ['gen_conv2d_converter']

QuantizedArgs = namedtuple("ConvQuantizedArgs",
                           "quantize_input in_signed in_width "
                           "wt_width quant_type "
                           "fake_bn wino_quantize")


def _per_sample_buffer(m, n):
    buf = getattr(m, "_fq_per_sample", None)
    if buf is None or buf.numel() != n or buf.device != m.current_input_max.device:
        buf = torch.empty(n, dtype=torch.float32, device=m.current_input_max.device)
        m._fq_per_sample = buf
    return buf


def _global_range(x, m, group):
    """Data parallel, range needed NOW (online quantisation): the shard's per-sample maxima are
    all-gathered; the Kahan mean over the global batch is then taken on every rank (by the quantiser itself,
    ops.forward_from_maxima).  Returns the gathered maxima [N]."""
    from ... import dist as fqdist
    per = _per_sample_buffer(m, x.shape[0])
    ops.input_range(x, cur_max=m.current_input_max, per_sample=per)
    m._fq_range_pending = False
    return fqdist.gather_per_sample(per, group)


def _input_path(x, m, lo_mode):
    """convert_conv2d.py:56-66.  Called directly when no gradient is needed, through _InputPath otherwise."""
    if torch.is_grad_enabled() and x.requires_grad:
        return _InputPath.apply(x, m, lo_mode)
    return _InputPath.forward(None, x, m, lo_mode)


def _weight_path(weight, bias, gamma, beta, mean, var, rows, bits):
    if torch.is_grad_enabled() and (weight.requires_grad or (gamma is not None and gamma.requires_grad)):
        return _WeightPath.apply(weight, bias, gamma, beta, mean, var, rows, bits)
    wq, bq, _ = ops.quant_weight(weight, rows, bits, gamma, beta, mean, var, bias)
    return wq, (bq if gamma is not None else bias)


def _is_wino(m):
    """convert_conv2d.py:71,80: per-channel quantisation of a 3x3 kernel in the Winograd domain."""
    qa = m.quantize_args
    return qa.wino_quantize != 'none' and qa.quant_type == 'channel' and tuple(m.kernel_size) == (3, 3)


def _wino_mats(m):
    """(G, pinv(G), pinv(G.T)) on the weight's device; the pseudo-inverses are the float32 ones NumPy computes on
    the host, as in the reference (:80-82), and are cached per block."""
    dev = m.weight.device
    mats = m.__dict__.get("_fq_wino")
    if mats is None or mats[0].device != dev:
        mats = tuple(torch.from_numpy(a).to(dev) for a in winograd_matrices(m.quantize_args.wino_quantize))
        m._fq_wino = mats
    return mats


class _WinoPath(torch.autograd.Function):
    """convert_conv2d.py:71-83: G w G^T -> per-channel fake-quant -> G+ . (G^T)+ ; the quantiser is straight
    through, the four matrix products get their exact adjoints."""

    @staticmethod
    def forward(ctx, weight, G, GI, GTI, bits):
        ctx.mats = (G, GI, GTI)
        return ops.quant_weight_wino(weight, G, GI, GTI, bits)[0]

    @staticmethod
    def backward(ctx, dwq):
        return ops.wino_backward(dwq.contiguous(), *ctx.mats), None, None, None, None


def _wino_path(weight, m):
    G, GI, GTI = _wino_mats(m)
    if torch.is_grad_enabled() and weight.requires_grad:
        return _WinoPath.apply(weight, G, GI, GTI, m.quantize_args.wt_width)
    return ops.quant_weight_wino(weight, G, GI, GTI, m.quantize_args.wt_width)[0]


def _planned_input_path(x, m, lo_mode, quantize, per):
    """ops.forward_online for block ``m`` through a cached C-side call plan (ops.InputPlan): the block's state tensors
    and quantiser settings are captured once per (mode, activation shape); a forward then costs one output
    allocation and one foreign call.  The plan is rebuilt when any tensor it borrowed has been replaced (``net.to()``
    re-packs the state arenas) -- identity checks on the block's own dicts, no Module.__getattr__."""
    d = m.__dict__
    bufs = m._buffers
    offline = quantize and m.quantize_input_offline
    cur = bufs["current_input_max"]
    qp = bufs["_fq_qparams"] if quantize else None
    key = (quantize, offline, lo_mode, ops.get_promotion(), x.shape)
    plans = d.get("_fq_in_plans")
    if plans is None:
        plans = d["_fq_in_plans"] = {}
    plan = plans.get(key)
    if (plan is None or plan.cur_max is not cur or plan.qparams is not qp or plan.per_sample is not per
            or plan.device != x.device or (offline and plan.input_max_ptr != m._parameters["input_max"].data_ptr())):
        qa = m.quantize_args
        if len(plans) >= 8:         # a loader with many ragged batch shapes: do not grow without bound
            plans.clear()
        plan = plans[key] = ops.InputPlan(x, qa.in_width, qa.in_signed, lo_mode,
                                          input_max=m._parameters["input_max"].data if offline else None,
                                          quantize=quantize, cur_max=cur, qparams=qp, per_sample=per)
    if x.dtype is not torch.float32:
        raise ops._ffi.FQError("x must be float32, got %s" % x.dtype)
    return plan.run(x)


class _InputPath(torch.autograd.Function):
    """convert_conv2d.py:56-66 on the device; backward = identity (ste_func.py:43-44)."""

    @staticmethod
    def forward(ctx, x, m, lo_mode):
        d = m.__dict__
        group = d.get("_fq_dist_group")
        if group is not None and not m.quantize_input_offline:
            qa = m.quantize_args
            allmax = _global_range(x, m, group)      # data parallel + online: range over the global batch first
            return ops.forward_from_maxima(x, allmax, qa.in_width, qa.in_signed, lo_mode, cur_max=m.current_input_max,
                                           qparams=m._fq_qparams)[0]
        # single process, or offline range: one fused launch.  Under data parallelism the per-sample maxima
        # are kept and the global mean is taken for ALL layers by one collective in update_ema().
        per = _per_sample_buffer(m, x.shape[0]) if group is not None else None
        y = _planned_input_path(x, m, lo_mode, True, per)
        d["_fq_range_pending"] = group is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy, None, None


def _range_only(x, m):
    """quantize_input switched off: the range is still tracked (convert_conv2d.py:55-57)."""
    d = m.__dict__
    group = d.get("_fq_dist_group")
    per = _per_sample_buffer(m, x.shape[0]) if group is not None else None
    _planned_input_path(x, m, ops.LO_ZERO, False, per)
    d["_fq_range_pending"] = group is not None


def _as_rows(tensors):
    """[len(tensors), n] view over the tensors when they are consecutive rows of one allocation, else None."""
    t0 = tensors[0]
    n = t0.numel()
    base = t0.untyped_storage().data_ptr()
    for i, t in enumerate(tensors):
        if (t.numel() != n or t.dtype != t0.dtype or not t.is_contiguous() or t.untyped_storage().data_ptr() != base
                or t.storage_offset() != t0.storage_offset() + i * n):
            return None
    return torch.empty(0, dtype=t0.dtype, device=t0.device).set_(t0.untyped_storage(), t0.storage_offset(),
                                                                  (len(tensors), n), (n, 1))


def pending_range_blocks(blocks):
    """Blocks whose current_input_max still holds the mean over this rank's shard only."""
    return [m for m in blocks if getattr(m, "_fq_range_pending", False)]


def local_range_arena(todo, into=None):
    """[L, N/R] per-sample maxima of the pending blocks as ONE tensor.  The blocks' buffers are rows of an arena
    (packed on first use, or moved into ``into`` when the caller owns the exchange buffer), so later steps
    need no packing kernel: the range kernels write straight into it."""
    rows = _as_rows([m._fq_per_sample for m in todo])
    if into is not None:
        if rows is None or rows.data_ptr() != into.data_ptr():
            for i, m in enumerate(todo):
                into[i].copy_(m._fq_per_sample)
                m._fq_per_sample = into[i]
        return into
    if rows is None:
        rows = torch.stack([m._fq_per_sample for m in todo])
        for i, m in enumerate(todo):
            m._fq_per_sample = rows[i]
    return rows


def finish_global_ranges(todo, allmax):
    """allmax: [R, L, N/R] per-sample maxima of every rank -> Kahan mean over the global batch in sample order,
    written to every block's current_input_max (one launch when update_ema() has packed them)."""
    world = allmax.shape[0]
    allmax = allmax.permute(1, 0, 2).reshape(len(todo), -1).contiguous()   # [L, N]
    cur = _as_rows([m.current_input_max for m in todo])
    if cur is not None:
        ops.mean_kahan(allmax, out=cur.view(-1))
    else:
        means = ops.mean_kahan(allmax)
        for i, m in enumerate(todo):
            m.current_input_max.copy_(means[i:i + 1])
    for m in todo:
        m._fq_range_pending = False
    return world


def sync_pending_ranges(blocks):
    """Data parallel: replace every block's shard-local current_input_max by the mean over the GLOBAL batch,
    with ONE all-gather of all layers' per-sample maxima and one batched Kahan-mean launch -- not a collective
    or a copy per layer.  (A dist.GradBucket built with ``net=`` goes further and lets the maxima ride in the tail
    of the gradient all-reduce.)"""
    from ... import dist as fqdist
    todo = pending_range_blocks(blocks)
    if not todo:
        return
    group = todo[0]._fq_dist_group
    local = local_range_arena(todo)                                       # [L, N/R]
    world = torch.distributed.get_world_size(group)
    allmax = fqdist.gather_per_sample(local.reshape(-1), group)           # [R, L, N/R]
    finish_global_ranges(todo, allmax.reshape(world, len(todo), -1))


class _WeightPath(torch.autograd.Function):
    """convert_conv2d.py:47-51 (fake-BN fold) + :70-95 (range, scale, fake-quant), fused.

    Backward: identity through the quantiser, then the fold's chain rule in the op order MXNet's
    autograd would replay ((W * gamma) / sd ; gamma * (b - mean) / sd + beta)."""

    @staticmethod
    def forward(ctx, weight, bias, gamma, beta, mean, var, rows, bits):
        fold = gamma is not None
        wq, bq, _ = ops.quant_weight(weight, rows, bits, gamma, beta, mean, var, bias)
        ctx.fold = fold
        ctx.has_bias = bias is not None
        if fold:
            ctx.save_for_backward(weight, bias if bias is not None else torch.zeros_like(gamma), gamma, mean, var)
            return wq, bq
        return wq, bias

    @staticmethod
    def backward(ctx, dwq, dbq):
        if not ctx.fold:
            return dwq, dbq, None, None, None, None, None, None
        weight, bias, gamma, mean, var = ctx.saved_tensors
        dweight, dgamma, dbias, dbeta = _fold_backward([dict(dwq=dwq, dbq=dbq, w=weight, gamma=gamma, mean=mean, var=var,
                                                             bias=bias if ctx.has_bias else None)])[0]
        return dweight, (dbias if ctx.has_bias else None), dgamma, dbeta, None, None, None, None


def _fold_backward(jobs):
    """(dweight, dgamma, dbias, dbeta) per job: one fused launch for all jobs that have a weight gradient
    (ops.fold_backward_multi), the op-by-op formula for the rest."""
    out = [None] * len(jobs)
    fused = [i for i, jb in enumerate(jobs) if jb["dwq"] is not None and jb["dwq"].is_cuda]
    if fused:
        for i, res in zip(fused, ops.fold_backward_multi([jobs[i] for i in fused])):
            out[i] = res
    for i, jb in enumerate(jobs):
        if out[i] is not None:
            continue
        weight, gamma, mean, var, bias, dwq, dbq = (jb[k] for k in ("w", "gamma", "mean", "var", "bias", "dwq", "dbq"))
        cout = weight.shape[0]
        sd = torch.sqrt(var + 1e-10)
        dweight = dgamma = dbias = dbeta = None
        if dwq is not None:
            da = dwq.reshape(cout, -1) / sd.reshape(-1, 1)
            dweight = (da * gamma.reshape(-1, 1)).reshape(weight.shape)
            dgamma = (da * weight.reshape(cout, -1)).sum(dim=1)
        if dbq is not None:
            dn = dbq / sd
            b = bias if bias is not None else torch.zeros_like(gamma)
            dgamma = dn * (b - mean) if dgamma is None else dgamma + dn * (b - mean)
            dbias = dn * gamma
            dbeta = dbq
        out[i] = (dweight, dgamma, dbias, dbeta)
    return out


class _MultiWeightPath(torch.autograd.Function):
    """The weight paths of every converted block in one launch per phase (ops.quant_weight_multi).
    Inputs per job: (weight, bias, gamma, beta) with None where absent; outputs per job: w_q (+ folded bias).
    Backward = _WeightPath.backward applied job by job."""

    @staticmethod
    def forward(ctx, plan, jobs, *tensors):
        ws, bs, _ = ops.quant_weight_multi(plan)
        ctx.plan = plan
        ctx.jobs = jobs
        ctx.tensors = tensors
        outs = []
        for i, jb in enumerate(jobs):
            outs.append(ws[i])
            if jb.get("gamma") is not None:
                outs.append(bs[i])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        out = [None, None]
        g = iter(grads)
        fold_jobs, slots = [], []
        for jb in ctx.jobs:
            dwq = next(g)
            if jb.get("gamma") is None:
                out.extend([dwq, None, None, None])
                continue
            dbq = next(g)
            fold_jobs.append(dict(dwq=dwq, dbq=dbq, w=jb["w"], gamma=jb["gamma"], mean=jb["mean"], var=jb["var"],
                                  bias=jb.get("bias")))
            slots.append(len(out))
            out.extend([None, None, None, None])
        # the fold backward of every block in ONE launch (the op-by-op formula is 6-8 launches per block)
        results = []
        if fold_jobs:
            if all(jb["dwq"] is not None and jb["dwq"].is_cuda for jb in fold_jobs):
                # the persistent tensors of these jobs are those of ctx.plan (rebuilt whenever one of them moves), so
                # their descriptors are built once and only the gradients' addresses change from step to step
                fplan = ctx.plan.__dict__.get("fold_plan")
                if fplan is None:
                    fplan = ctx.plan.__dict__["fold_plan"] = ops.FoldBackwardPlan(fold_jobs)
                results = fplan.run([jb["dwq"] for jb in fold_jobs], [jb["dbq"] for jb in fold_jobs])
            else:
                results = _fold_backward(fold_jobs)
        for pos, jb, (dweight, dgamma, dbias, dbeta) in zip(slots, fold_jobs, results):
            out[pos:pos + 4] = [dweight, dbias if jb["bias"] is not None else None, dgamma, dbeta]
        return tuple(out)


_SIG_TENSORS = ("weight", "bias", "gamma", "beta", "running_mean", "running_var")


def _weights_signature(blocks):
    """What decides the job list of :func:`prequantize_weights` and whether its cached C job table is still valid:
    per block the tri-state / switches and the storage address of EVERY tensor a job borrows (a cached table holds raw
    pointers: it is stale as soon as one of them is replaced).  Reads the blocks' own dicts only."""
    sig = []
    for m in blocks:
        d, p = m.__dict__, m._parameters
        sig.append(d.get("fixed_params"))
        sig.append(d.get("enable_quantize"))
        sig.append(id(d.get("quantize_args")))
        for name in _SIG_TENSORS:
            t = p.get(name)
            sig.append(None if t is None else t.data_ptr())
    return sig


def _collect_weight_jobs(blocks):
    jobs, owners = [], []
    for m in blocks:
        qa = m.quantize_args
        if isinstance(m, nn.Conv2d):
            if m.fixed_params == 1:
                continue
            fold = qa.fake_bn
            if m.enable_quantize:
                if _is_wino(m):
                    continue                  # Winograd-domain blocks take their own per-block path
                bits, rows = qa.wt_width, _weight_rows(m)
            elif fold:
                bits, rows = 0, 1
            else:
                continue
            jb = {"w": m.weight, "rows": rows, "bits": bits}
            if fold:
                jb.update(gamma=m.gamma, beta=m.beta, mean=m.running_mean, var=m.running_var, bias=m.bias)
        elif isinstance(m, nn.Linear):
            if not m.enable_quantize:
                continue
            jb = {"w": m.weight, "rows": m.out_features if qa.quant_type == 'channel' else 1, "bits": qa.wt_width}
        else:
            continue
        if not m.weight.is_cuda:
            return [], []
        if any(t is not None and (t.dtype != torch.float32 or not t.is_contiguous()) for t in jb.values()
               if isinstance(t, torch.Tensor)):
            continue                          # e.g. a channels_last weight: the per-block path copies it afresh each call
        jobs.append(jb)
        owners.append(m)
    return jobs, owners


def prequantize_weights(net, blocks):
    """Net-level forward pre-hook body: run the weight path of every block that needs one in this forward as a
    single multi-tensor launch and hand each block its result through ``_fq_pre``.

    Host cost matters here (an eager forward of a small network is bound by it): the job list, the C job table and
    -- without autograd -- the output buffers are cached on the net and revalidated per forward by one flat list
    comparison (:func:`_weights_signature`)."""
    nd = net.__dict__
    sig = _weights_signature(blocks)
    cache = nd.get("_fq_weight_plan")
    if cache is None or cache["sig"] != sig:
        jobs, owners = _collect_weight_jobs(blocks)
        # the plan holds raw DLTensors: it is only valid while EVERY tensor of every job is the same storage
        cache = nd["_fq_weight_plan"] = {
            "sig": sig, "jobs": jobs, "owners": owners, "plan": ops.WeightPlan(jobs) if len(jobs) >= 2 else None,
            "folds": [jb.get("gamma") is not None for jb in jobs], "outs": {},
            "grad_tensors": [t for jb in jobs for t in (jb["w"], jb.get("gamma"), jb.get("beta"), jb.get("bias"))
                             if t is not None]}
    plan = cache["plan"]
    if plan is None:
        return                                # nothing to batch: the per-block path is just as good
    jobs, owners = cache["jobs"], cache["owners"]
    if torch.is_grad_enabled() and any(t.requires_grad for t in cache["grad_tensors"]):
        flat = []
        for jb in jobs:
            flat.extend([jb["w"], jb.get("bias") if jb.get("gamma") is not None else None, jb.get("gamma"), jb.get("beta")])
        outs = iter(_MultiWeightPath.apply(plan, jobs, *flat))
        for m, fold in zip(owners, cache["folds"]):
            wq = next(outs)
            m.__dict__["_fq_pre"] = (wq, next(outs) if fold else None)
        return
    # no autograd: nothing outlives the forward, so the quantised weights are written into buffers that persist per
    # stream (stream order protects the previous forward's readers) -- no allocation and no 3 x len(jobs) view
    # objects per call.  Not while a CUDA graph is being captured: a graph gets buffers of its own pool.
    if torch.cuda.is_current_stream_capturing():
        ws, bs, _ = ops.quant_weight_multi(plan)
        for i, m in enumerate(owners):
            m.__dict__["_fq_pre"] = (ws[i], bs[i])
        return
    raw = ops._ffi._raw_stream(plan.device.index or 0)
    held = cache["outs"].get(raw)
    if held is None:
        if len(cache["outs"]) >= 4:
            cache["outs"].clear()
        bufs = ops.weight_multi_buffers(plan)
        held = cache["outs"][raw] = (bufs, [(w, b) for w, b in zip(bufs[3], bufs[4])])
    ops.quant_weight_multi(plan, held[0])
    for m, pre in zip(owners, held[1]):
        m.__dict__["_fq_pre"] = pre


def _weight_rows(m):
    qt = m.quantize_args.quant_type
    if qt == 'channel':
        return m.out_channels
    if qt == 'group':
        return m.groups
    return 1


def _conv2d_forward(self, x):
    qa = self.quantize_args
    weight, bias = self.weight, self.bias
    fold = self.fixed_params != 1 and qa.fake_bn
    if self.enable_quantize:
        # Quantize input (convert_conv2d.py:55-66)
        if qa.quantize_input:
            if self.quantize_input:
                lo_mode = ops.LO_NEG_MAX if qa.in_signed else ops.LO_ZERO
                x = _input_path(x, self, lo_mode)
            else:       # the range is still tracked (convert_conv2d.py:56)
                _range_only(x.detach(), self)
        # Simulate quantization for weight (:69-97)
        pre = self.__dict__.pop("_fq_pre", None)       # set by the net-level multi-tensor launch, used once
        if pre is not None:
            weight_q, bias = pre[0], (pre[1] if fold else bias)
        elif self.fixed_params != 1 and _is_wino(self):
            if fold:        # :47-51 first, then the Winograd-domain quantiser on the folded weight
                weight, bias = _weight_path(weight, bias, self.gamma, self.beta, self.running_mean,
                                            self.running_var, 1, 0)
            weight_q = _wino_path(weight, self)
        elif self.fixed_params != 1:
            if fold:
                weight_q, bias = _weight_path(weight, bias, self.gamma, self.beta, self.running_mean,
                                              self.running_var, _weight_rows(self), qa.wt_width)
            else:
                weight_q, bias = _weight_path(weight, bias, None, None, None, None, _weight_rows(self), qa.wt_width)
        else:
            weight_q = weight
    else:
        pre = self.__dict__.pop("_fq_pre", None)
        if pre is not None:
            weight_q, bias = pre
        elif fold:       # fold only (:47-51 runs even when quantisation is disabled)
            weight_q, bias = _weight_path(weight, bias, self.gamma, self.beta, self.running_mean,
                                          self.running_var, 1, 0)
        else:
            weight_q = weight

    if self.fixed_params == 0:      # :101-105
        self.fixed_params = 1
        with torch.no_grad():
            self.weight.copy_(weight_q)
            if bias is not None:
                self.bias.copy_(bias)

    # Normal convolution
    return self.origin_forward(x, weight_q, bias)


def _add_quantize_input_params(m):
    m.quantize_input_offline = False
    dev = m.weight.device
    # non-persistent buffers follow .cuda()/.to() but stay out of the state dict, like the reference's
    # plain attribute (current_input_max is not checkpointed)
    m.register_buffer("current_input_max", torch.zeros(1, dtype=torch.float32, device=dev), persistent=False)
    m.register_parameter("input_max", nn.Parameter(torch.zeros(1, dtype=torch.float32, device=dev),
                                                   requires_grad=False))
    m.register_buffer("_fq_qparams", torch.zeros(4, dtype=torch.float32, device=dev), persistent=False)


def _add_fake_bn_params(m):
    c, dev = m.out_channels, m.weight.device
    m.register_parameter("gamma", nn.Parameter(torch.ones(c, device=dev)))
    m.register_parameter("beta", nn.Parameter(torch.zeros(c, device=dev)))
    m.register_parameter("running_mean", nn.Parameter(torch.zeros(c, device=dev), requires_grad=False))
    m.register_parameter("running_var", nn.Parameter(torch.ones(c, device=dev), requires_grad=False))
    # plain attributes in the reference (set by the pre-hook, :151-153); buffers here so that they follow the net
    m.register_buffer("current_mean", torch.zeros(c, device=dev), persistent=False)
    m.register_buffer("current_var", torch.zeros(c, device=dev), persistent=False)


def _add_fake_bn_ema_hook(m):
    """convert_conv2d.py:144-154: batch statistics of the RAW convolution output for the EMA."""
    def _ema_hook(m, x):
        with torch.no_grad():
            y = m.origin_forward(x[0], m.weight, m.bias)        # the reference's second convolution (:149)
            # mean = y.sum(axis=(0,2,3)) / num ; var = ((y - mean) ** 2).sum(axis=(0,2,3)) / num   (:150-153)
            # one pass over y; written in place so that the net's packed state (and a captured graph) sees it
            group = getattr(m, "_fq_dist_group", None)
            parts = m.__dict__.get("_fq_stats_parts") if group is not None else None
            ops.channel_stats(y, mean=m.current_mean, var=m.current_var, parts=parts)
            m._fq_stats_pending = parts is not None             # shard-local until update_ema() combines the ranks
    m.register_forward_pre_hook(_ema_hook)


def sync_pending_stats(blocks, arenas):
    """Data parallel: the fake-BN batch statistics of every block become those of the GLOBAL batch -- ONE all-gather
    of all layers' float64 {n, S1, S2, K} records and one finish launch (the reference's
    current_mean/current_var, convert_conv2d.py:150-153, over N = sum of the ranks' shards)."""
    todo = [m for m in blocks if getattr(m, "_fq_stats_pending", False)]
    rec = arenas.get("stats_parts")
    if not todo or rec is None:
        return
    group = todo[0]._fq_dist_group
    parts = rec["parts"]
    world = torch.distributed.get_world_size(group)
    gathered = rec.get("gathered")
    if gathered is None or gathered.shape[0] != world or gathered.device != parts.device:
        gathered = rec["gathered"] = torch.empty((world,) + tuple(parts.shape), dtype=parts.dtype, device=parts.device)
    torch.distributed.all_gather_into_tensor(gathered.view(-1), parts.view(-1), group=group)
    ops.channel_stats_finish(gathered, mean=arenas["running_mean"]["current"], var=arenas["running_var"]["current"])
    for m in todo:
        m._fq_stats_pending = False


def gen_conv2d_converter(weight_width=8, quant_type="layer",
                         quantize_input=True, input_signed=False, input_width=8,
                         fake_bn=False, wino_quantize="none"):
    assert wino_quantize in ("none", "F23", "F43", "F63")

    def _converter(m):
        assert isinstance(m, nn.Conv2d)

        if quantize_input:
            _add_quantize_input_params(m)
        if fake_bn:
            _add_fake_bn_params(m)
            _add_fake_bn_ema_hook(m)
        m.origin_forward = m._conv_forward
        m.forward = types.MethodType(_conv2d_forward, m)
        m.quantize_args = QuantizedArgs(in_signed=input_signed, in_width=input_width, wt_width=weight_width,
                                        quantize_input=quantize_input, fake_bn=fake_bn, quant_type=quant_type,
                                        wino_quantize=wino_quantize)
        m.fixed_params = -1
        m.enable_quantize = True
        m.quantize_input = quantize_input
    return _converter
