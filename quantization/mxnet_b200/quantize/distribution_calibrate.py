"""KL calibration (quantize/distribution_calibrate.py of the reference), device resident.

The reference copies every hooked layer input to the host (``x.asnumpy()``), histograms it with six
single-threaded NumPy passes and searches the threshold in a pure-Python O(bins^2) loop.  Here the
forward hook launches one histogram kernel on the activation where it lives (4 B/element, read
once), per-batch counts of ALL layers are folded into the float32 histograms by one launch, and the
KL search runs one CUDA block per candidate threshold.  With ``torch.distributed`` initialised the
batch is sharded across ranks: first-batch maxima are max-all-reduced and the integer counts of up to
``ring_slots`` batches are sum-all-reduced together and then folded in batch order, so every rank ends with
the bit-identical histograms of a single-GPU run over the global batches.

Reference behaviours kept: the first batch's max is frozen for all later batches (:97-101); zeros
are ignored (:40); float32 accumulation in batch order (:47,:103-104); the 2049th bin when
``max_ >= 256`` (and the ValueError when only some batches have it); the asserts of :35-36, evaluated on
EVERY batch by a flag the histogram kernel raises (negative value, NaN, non-positive max) and reported once
at the end.  The hooked inputs are histogrammed after the forward, not copied at hook time; an input that is
modified in place later in the same forward is detected (tensor version counter) and raises.
"""
import warnings

import numpy as np
import torch
from tqdm import tqdm

from .. import dist as fqdist
from .. import ops

__all__ = ['collect_feature_maps', 'kl_calibrate', 'kl_calibrate_all']


class _Collector(dict):
    """dict(block -> numpy value) that also keeps the device-resident stack (``.device``) so that a
    following :func:`kl_calibrate_all` does not upload anything."""
    device = None
    order = ()


def collect_feature_maps(net, bins, loader, ctx=None, tqdm_desc="Collect FM", group=None, ring_slots=8):
    """
    Collect feature maps and record discrete histograms.
    :param net: converted torch.nn.Module
    :param bins: int
        Number of bins to generate discrete histograms.
    :param loader: iterable of (X, y) batches
    :param ctx: torch.device (or None: keep X where it is)
    :param group, ring_slots: data-parallel runs only -- process group, and how many batches share one
        sum-all-reduce of the integer counts (results do not depend on it)
    :return: (hist_collector, fm_max_collector) keyed by block, values numpy float32 histogram /
        numpy.float32 max, as in the reference.
    """
    quantized_blocks = net.collect_quantized_blocks()
    n_blk = len(quantized_blocks)
    index = {id(b): i for i, b in enumerate(quantized_blocks)}
    group = fqdist.active_group(group)

    state = {}      # allocated on the first hooked tensor's device

    def _alloc(dev):
        state["hist"] = torch.zeros(n_blk, bins + 1, dtype=torch.float32, device=dev)
        state["minmax"] = torch.zeros(n_blk, 2, dtype=torch.float32, device=dev)
        state["bad"] = torch.zeros(n_blk, dtype=torch.int32, device=dev)      # asserts of :35-36, every batch
        state["seen_hist"] = []
        state["dev"] = dev

    def _alloc_ring(max_numel):
        # per-block flag "this batch produced a 2049th bin" (deferred length check), taken from the GLOBAL counts.
        # Data parallel: 32-bit counts on the wire while no bin of a global batch can reach 2^32 (max_numel bounds
        # one rank's count per bin; the hook re-checks every later input against it).
        state["max_numel"] = max_numel if group is not None else None
        state["ring"] = fqdist.CountsRing(
            n_blk, bins + 1, state["dev"], group=group, slots=ring_slots, max_count=state["max_numel"],
            accumulate=lambda c, first: ops.hist_accumulate(c.reshape(-1), state["hist"].view(-1), first, None),
            on_reduced=lambda c: state["seen_hist"].append((c[:, :, bins] != 0).to(torch.int32)))
        state["ring"].prime()        # data parallel: connect NCCL for this message size outside the batch loop

    """ Add hooks to quantized blocks """
    hooks = []
    first_batch = {}        # block index -> inputs of the first batch the block runs in (kept until its max is known)
    have_max = set()        # blocks whose max is frozen
    called = set()
    called_rows = []        # per batch: which blocks were called (the 2049th-bin check only looks at those)
    pending = {}            # block index -> this batch's input, for the single multi-tensor launch
    versions = {}           # id(tensor) -> version at hook time: an in-place change before the launch is an error
    n_batches = 0

    def _collect(m, x, y):
        x = x[0].detach()
        if not state:
            _alloc(x.device)
        i = index[id(m)]
        called.add(i)
        batch_called.add(i)
        if state.get("max_numel") is not None and x.numel() > state["max_numel"] and \
                state["ring"].dtype == torch.int32:
            raise RuntimeError("a layer input of %d elements is larger than any of the first batch (%d): the 32-bit "
                               "count exchange was sized for the first batch" % (x.numel(), state["max_numel"]))
        if i not in have_max:
            # the first batch in which THIS block runs fixes its max (fm_max_collector.get(m) is None, :97-102) --
            # batch 0 for a plain network, a later one for a block behind a data-dependent branch
            first_batch.setdefault(i, []).append(x)
            versions[id(x)] = x._version
        elif x.data_ptr() % 16 == 0 and i not in pending:
            pending[i] = x          # histogrammed together with the other layers after the forward
            versions[id(x)] = x._version
        else:
            ops.hist_nonzero(x, state["minmax"][i, 1:2], bins, state["ring"].slot()[i], bad_flag=state["bad"][i:i + 1])

    def _unchanged(x):
        if versions.get(id(x), x._version) != x._version:
            raise RuntimeError("a hooked layer input was modified in place after its block ran and before its "
                               "histogram was taken (the reference copies it at hook time, "
                               "distribution_calibrate.py:80); make that op out-of-place for calibration")
        return x
    batch_called = set()
    for blk in quantized_blocks:
        hooks.append(blk.register_forward_hook(_collect))

    """ Collect feature maps """
    try:
        with tqdm(total=len(loader), desc=tqdm_desc, disable=None) as pbar, torch.no_grad():
            for X in _prefetch(loader, ctx):
                _ = net(X)
                if first_batch:
                    # First chunk of these blocks: min/max of everything the block saw, then its histogram
                    for i, xs in first_batch.items():
                        mm = state["minmax"][i]
                        ops.minmax(_unchanged(xs[0]), out=mm)
                        for extra in xs[1:]:
                            mm2 = ops.minmax(_unchanged(extra))
                            mm[0:1].copy_(torch.minimum(mm[0:1], mm2[0:1]))
                            mm[1:2].copy_(torch.maximum(mm[1:2], mm2[1:2]))
                    fqdist.sync_first_batch_minmax(state["minmax"], group)
                    if "ring" not in state:
                        _alloc_ring(max(x.numel() for xs in first_batch.values() for x in xs))
                    for i, xs in first_batch.items():
                        for x in xs:
                            ops.hist_nonzero(x, state["minmax"][i, 1:2], bins, state["ring"].slot()[i],
                                             bad_flag=state["bad"][i:i + 1])
                    have_max.update(first_batch)
                    first_batch.clear()
                if pending:
                    # every layer of the batch in ONE launch, blocks shared out by tensor size
                    order = sorted(pending)
                    if order == list(range(n_blk)):
                        ops.hist_nonzero_multi([_unchanged(pending[i]) for i in order], state["minmax"], 2, 1, bins,
                                               state["ring"].slot(), bad_flags=state["bad"])
                    else:
                        for i in order:
                            ops.hist_nonzero(_unchanged(pending[i]), state["minmax"][i, 1:2], bins,
                                             state["ring"].slot()[i], bad_flag=state["bad"][i:i + 1])
                    pending.clear()
                versions.clear()
                if state:
                    # hist_collector[m] = last_hist + hist.astype(float32), all blocks in one launch; with
                    # ranks, one sum-all-reduce per `ring_slots` batches and the adds replayed in batch order
                    state["ring"].commit()
                    row = np.zeros(n_blk, bool)
                    row[list(batch_called)] = True
                    called_rows.append(row)
                batch_called.clear()
                n_batches += 1
                pbar.update(1)
            if state:
                state["ring"].flush()
    finally:
        """ Delete hooks """
        for h in hooks:
            h.remove()

    hist_collector, fm_max_collector = _Collector(), _Collector()
    if not state:
        return hist_collector, fm_max_collector

    # one device->host transfer for everything, then the reference's deferred checks
    hist = state["hist"].cpu().numpy()
    minmax = state["minmax"].cpu().numpy()
    seen = torch.cat(state["seen_hist"]).cpu().numpy() if state["seen_hist"] else np.zeros((0, n_blk), np.int32)
    bad = state["bad"].cpu().numpy()
    if group is not None:       # an assert that fires on any rank fires on all
        t = state["bad"].clone()
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX, group=group)
        bad = t.cpu().numpy()
    was_called = np.stack(called_rows) if called_rows else np.zeros((0, n_blk), bool)
    for i, m in enumerate(quantized_blocks):
        if i not in called:
            continue
        assert minmax[i, 0] >= 0., "Activation should >=0"
        assert minmax[i, 1] > 0, "Bad distribution: all zero-value"
        # the same two asserts on the later batches (and NaN anywhere), from the kernels' device flags
        assert not bad[i], "Activation should >=0"
        col = seen[:len(was_called), i][was_called[:len(seen), i]]     # batches in which this block ran
        if col.any() and not col.all():
            # np.bincount gave 2049 bins for some batches and 2048 for others: `last_hist + hist` raises
            raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,)" % (bins, bins + 1))
        n = bins + 1 if col.any() else bins
        hist_collector[m] = hist[i, :n].copy()
        fm_max_collector[m] = np.float32(minmax[i, 1])
    hist_collector.device = state["hist"]
    hist_collector.order = tuple(quantized_blocks)
    fm_max_collector.device = state["minmax"][:, 1].contiguous()
    fm_max_collector.order = tuple(quantized_blocks)
    return hist_collector, fm_max_collector


def _prefetch(loader, ctx):
    """Yield each batch on ``ctx``: batch k+1 is copied host->device on a side stream while the network runs on
    batch k (pinned host memory makes the copy asynchronous).  Two persistent device buffers are used in turn;
    before one is overwritten the host waits for the step that read it, which also keeps the host at most two
    steps ahead of the GPU -- no allocation per batch, no unbounded queue of copies."""
    if ctx is None or torch.device(ctx).type != "cuda":
        for X, _ in loader:
            yield X if ctx is None else X.to(ctx)
        return
    dev = torch.device(ctx)
    copy_stream = torch.cuda.Stream(device=dev)
    it = iter(loader)
    bufs, consumed = [None, None], [None, None]
    fetched = 0

    def fetch():
        nonlocal fetched
        try:
            X, _ = next(it)
        except StopIteration:
            return None
        if X.device == dev:
            return X, None, None
        slot = fetched % 2
        fetched += 1
        if consumed[slot] is not None:
            consumed[slot].synchronize()        # the step that read this buffer two batches ago has finished
        if bufs[slot] is None or bufs[slot].shape != X.shape or bufs[slot].dtype != X.dtype or \
                bufs[slot].stride() != X.stride():
            bufs[slot] = torch.empty_like(X, device=dev)       # keeps the batch's memory format (channels_last)
            copy_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(copy_stream):
            bufs[slot].copy_(X, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return bufs[slot], ev, slot
    nxt = fetch()
    while nxt is not None:
        Xd, ev, slot = nxt
        if ev is not None:
            torch.cuda.current_stream(dev).wait_event(ev)
        nxt = fetch()           # start the next copy before this batch's forward is queued
        yield Xd
        if slot is not None:    # the consumer has queued everything that reads Xd
            consumed[slot] = torch.cuda.Event()
            consumed[slot].record(torch.cuda.current_stream(dev))


def kl_calibrate(data, levels, min_bins, bins):
    """
    KL-divergence calibration for offline-quantization (one histogram).
    :param data: numpy.ndarray or torch.Tensor, discrete histogram (length bins or bins + 1)
    :return: int, best threshold bin.
    """
    assert min_bins >= levels, f"min_bins should be greater than levels ({min_bins} vs. {levels})"
    if isinstance(data, np.ndarray):
        data = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).cuda()
    margin = torch.empty(1, dtype=torch.float64, device=data.device)
    best, _ = ops.kl_search(data.reshape(1, -1), levels, min_bins, bins, margin=margin)
    _flag_near_ties(best, margin)
    return int(best[0])


def _flag_near_ties(best, margin):
    """Warn when a chosen bin wins by less than ops.KL_TIE_MARGIN (relative): see fq.h fq_kl_search."""
    m = margin.cpu().numpy()
    tied = np.flatnonzero(m < ops.KL_TIE_MARGIN)
    if tied.size:
        b = best.cpu().numpy()
        warnings.warn("KL threshold search: near-tie between candidate bins for layer(s) %s (best bin(s) %s, relative "
                      "margin(s) %s): the reference's own choice here depends on its math library's rounding"
                      % (tied.tolist(), b[tied].tolist(), m[tied].tolist()), RuntimeWarning, stacklevel=3)
    return m


def kl_calibrate_all(hists, levels, min_bins, bins, fm_max=None, check_ties=True):
    """All layers at once: ``hists`` is a [layers, n] tensor (or a hist_collector from
    :func:`collect_feature_maps`).  Returns int32 best bins on the device and, when ``fm_max`` is
    given, the float32 thresholds ``(best + 0.5) * (fm_max / bins)`` (simulate_quantization.py:310).
    ``check_ties`` reads back one float64 per layer and warns about near-ties (ops.KL_TIE_MARGIN)."""
    assert min_bins >= levels, f"min_bins should be greater than levels ({min_bins} vs. {levels})"
    if isinstance(hists, _Collector):
        dev_h = hists.device
        lens = {len(v) for v in hists.values()}
        # a stack can only be searched in one launch when every histogram has the same length
        if len(lens) == 1:
            n = lens.pop()
            hists = dev_h[:, :n].contiguous()
        else:
            # blocks the calibration never reached have no histogram: their row keeps best = min_bins
            outs = [kl_calibrate_all(dev_h[i:i + 1, :len(hists[m])].contiguous(), levels, min_bins, bins,
                                     check_ties=check_ties)
                    if m in hists else torch.full((1,), min_bins, dtype=torch.int32, device=dev_h.device)
                    for i, m in enumerate(hists.order)]
            best = torch.cat(outs)
            return (best, ops.kl_threshold(best, _max_of(fm_max), bins)) if fm_max is not None else best
    margin = torch.empty(1 if hists.dim() == 1 else hists.shape[0], dtype=torch.float64, device=hists.device) \
        if check_ties else None
    best, _ = ops.kl_search(hists, levels, min_bins, bins, margin=margin)
    if check_ties:
        _flag_near_ties(best, margin)
    if fm_max is None:
        return best
    return best, ops.kl_threshold(best, _max_of(fm_max), bins)


def _max_of(fm_max):
    if isinstance(fm_max, _Collector):
        return fm_max.device
    return fm_max
