"""Functional wrappers: torch CUDA tensors in, libfq_b200 kernels launched on torch's current stream.

Each function corresponds to one C-ABI entry point of include/fq.h and cites the reference code it
replaces.  Nothing here synchronises the host or touches the CPU.
"""
import torch

from . import _ffi
from ._ffi import LO_NEG_MAX, LO_ZERO, STE_CLIP_MASK, STE_IDENTITY, check_call, current_stream, dl, ptr, workspace

_PROMOTIONS = {"legacy": _ffi.PROMOTION_LEGACY, "nep50": _ffi.PROMOTION_NEP50}
_promotion = "legacy"


def set_promotion(mode):
    """How the reference's host scalar math is replayed: "legacy" (NumPy 1.x, what an MXNet 1.x
    install computes; default) or "nep50" (NumPy >= 2)."""
    global _promotion
    if mode not in _PROMOTIONS:
        raise ValueError("promotion must be 'legacy' or 'nep50'")
    _promotion = mode


def get_promotion():
    return _promotion


def _promo(p):
    return _PROMOTIONS[_promotion if p is None else p]


def _f32(x, name="x"):
    if x.dtype != torch.float32:
        raise _ffi.FQError("%s must be float32, got %s" % (name, x.dtype))
    return x.contiguous()


def _act(x, name="x"):
    """An activation as the kernels may read it: float32 and dense.  A ``channels_last`` tensor is dense memory in
    N, H, W, C order; for everything that is elementwise, or a reduction over whole samples, or over the whole tensor
    (quantisers, per-sample / global ranges, histograms) the order inside a sample does not matter, so its memory is
    handed over as it lies -- as the contiguous NHWC view -- instead of through an NCHW copy (two extra passes)."""
    if x.dtype != torch.float32:
        raise _ffi.FQError("%s must be float32, got %s" % (name, x.dtype))
    if x.is_contiguous():
        return x
    if x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last):
        return x.permute(0, 2, 3, 1)
    return x.contiguous()


def _ew(x, out, codes_dtype=None):
    """(input as the kernel reads it, output to return, output as the kernel writes it) for an elementwise kernel: a
    channels_last activation is read and written in place of its memory order (the result is channels_last too);
    with a codes output, or any other non-contiguous layout, the input is copied to NCHW first as before."""
    xa = _act(x)
    aliased = xa is not x and xa.data_ptr() == x.data_ptr()
    if aliased and codes_dtype is None and (out is None or out.is_contiguous(memory_format=torch.channels_last)):
        out = torch.empty_like(x) if out is None else out           # preserve_format: channels_last like x
        return xa, out, out.permute(0, 2, 3, 1)
    if aliased:
        xa = x.contiguous()
    out = torch.empty_like(xa) if out is None else out
    return xa, out, out


def _lib():
    return _ffi.load()


# ---- K1 ------------------------------------------------------------------------------------------
def absmax_rows(x, rows, out=None):
    """max |x| per row of x.view(rows, -1).  convert_conv2d.py:56,75,86,92."""
    x = _f32(x)
    out = torch.empty(rows, dtype=torch.float32, device=x.device) if out is None else out
    a, o = dl(x), dl(out)
    check_call(_lib().fq_absmax_rows(a.ptr, rows, o.ptr, workspace(x.device), current_stream()))
    return out


def minmax(x, out=None):
    """{min, max} of x.  nn/quantized_conv.py:68-69; distribution_calibrate.py:34-35."""
    x = _act(x)
    out = torch.empty(2, dtype=torch.float32, device=x.device) if out is None else out
    a, o = dl(x), dl(out)
    check_call(_lib().fq_minmax(a.ptr, o.ptr, workspace(x.device), current_stream()))
    return out


def mean_kahan(v, out=None):
    """MXNet CPU ``mean`` of a vector."""
    v = _f32(v)
    rows = 1 if v.dim() < 2 else v.numel() // v.shape[-1]
    out = torch.empty(rows, dtype=torch.float32, device=v.device) if out is None else out
    a, o = dl(v), dl(out)
    check_call(_lib().fq_mean_kahan(a.ptr, o.ptr, current_stream()))
    return out


def input_range(x, n_samples=None, cur_max=None, per_sample=None):
    """current_input_max = mean_n max_chw |x| in one launch.  convert_conv2d.py:56."""
    x = _act(x)
    n_samples = x.shape[0] if n_samples is None else n_samples
    cur_max = torch.empty(1, dtype=torch.float32, device=x.device) if cur_max is None else cur_max
    a, c, p = dl(x), dl(cur_max), dl(per_sample)
    check_call(_lib().fq_input_range(a.ptr, n_samples, ptr(p), c.ptr, workspace(x.device), current_stream()))
    return cur_max


def channel_stats(y, mean=None, var=None, parts=None, finish=True):
    """Per-channel batch mean and (biased) variance of an [N, C, H, W] tensor for the fake-BN EMA
    (convert_conv2d.py:148-153) in ONE pass over y (4 B/element instead of the six passes of the op-by-op formula).
    ``parts`` (float64 [C, 4]) also receives the {n, S1, S2, K} records that :func:`channel_stats_finish` combines
    across ranks; ``finish=False`` writes the records only."""
    y = _f32(y, "y")
    c = y.shape[1]
    if finish:
        mean = torch.empty(c, dtype=torch.float32, device=y.device) if mean is None else mean
        var = torch.empty(c, dtype=torch.float32, device=y.device) if var is None else var
    else:
        mean = var = None
    a, m, v, p = dl(y), dl(mean), dl(var), dl(parts)
    check_call(_lib().fq_channel_stats(a.ptr, ptr(m), ptr(v), ptr(p), workspace(y.device), current_stream()))
    return mean, var


def channel_stats_finish(parts, mean=None, var=None):
    """parts: float64 [R, C, 4] records of R ranks -> (mean, var) float32 [C] of the global batch."""
    c = parts.shape[-2]
    mean = torch.empty(c, dtype=torch.float32, device=parts.device) if mean is None else mean
    var = torch.empty(c, dtype=torch.float32, device=parts.device) if var is None else var
    p, m, v = dl(parts), dl(mean), dl(var)
    check_call(_lib().fq_channel_stats_finish(p.ptr, m.ptr, v.ptr, current_stream()))
    return mean, var


def scale_from_max(max_, bits, signed, lo_mode, qparams=None, promotion=None):
    """{d, s, lo, hi} from a device-resident range.  convert_conv2d.py:57-64 + ste_func.py:41."""
    qparams = torch.empty(4, dtype=torch.float32, device=max_.device) if qparams is None else qparams
    m, q = dl(_f32(max_, "max_")), dl(qparams)
    check_call(_lib().fq_scale_from_max(m.ptr, bits, int(bool(signed)), lo_mode, _promo(promotion), q.ptr,
                                        current_stream()))
    return qparams


# ---- K2 ------------------------------------------------------------------------------------------
def _codes_like(x, codes_dtype):
    return None if codes_dtype is None else torch.empty(x.shape, dtype=codes_dtype, device=x.device)


def forward_scalar(x, qparams, out=None, codes_dtype=None):
    """y = roundf(clip(x, lo, hi) / d) * s with device-resident {d, s, lo, hi}.  ste_func.py:41."""
    x, out, ov = _ew(x, out, codes_dtype)
    codes = _codes_like(x, codes_dtype)
    a, q, o, c = dl(x), dl(qparams), dl(ov), dl(codes)
    check_call(_lib().fq_forward_scalar(a.ptr, q.ptr, o.ptr, ptr(c), current_stream()))
    return (out, codes) if codes_dtype is not None else out


def forward_scalar_host(x, d, s, lo=0.0, hi=0.0, clip=True, out=None, codes_dtype=None):
    """Same with host scalars; clip=False is ste_func.py:39."""
    x, out, ov = _ew(x, out, codes_dtype)
    codes = _codes_like(x, codes_dtype)
    a, o, c = dl(x), dl(ov), dl(codes)
    check_call(_lib().fq_forward_scalar_host(a.ptr, d, s, lo, hi, int(bool(clip)), o.ptr, ptr(c), current_stream()))
    return (out, codes) if codes_dtype is not None else out


def forward_rows(x, scale, out=None, codes_dtype=None):
    """y[r] = roundf(x[r] / (scale[r] + 1e-10)) * scale[r].  ste_func.py:39 with an NDArray scale."""
    x = _f32(x)
    scale = _f32(scale, "scale").reshape(-1)
    out = torch.empty_like(x) if out is None else out
    codes = _codes_like(x, codes_dtype)
    a, s, o, c = dl(x), dl(scale), dl(out), dl(codes)
    check_call(_lib().fq_forward_rows(a.ptr, scale.numel(), s.ptr, o.ptr, ptr(c), current_stream()))
    return (out, codes) if codes_dtype is not None else out


def forward_online(x, bits=8, signed=False, lo_mode=LO_ZERO, input_max=None, quantize=True, n_samples=None,
                   out=None, cur_max=None, qparams=None, per_sample=None, codes_dtype=None, promotion=None):
    """The whole input path of a converted block without a host round trip (convert_conv2d.py:56-66):
    range launch + streaming quantiser online, one fused launch when the range comes from ``input_max``.

    Returns (y, cur_max, qparams[, codes]); y is None when ``quantize`` is False (range tracking only).
    """
    y = yv = None
    if quantize:
        x, y, yv = _ew(x, out, codes_dtype)
    else:
        x = _act(x)
    n_samples = x.shape[0] if n_samples is None else n_samples
    cur_max = torch.empty(1, dtype=torch.float32, device=x.device) if cur_max is None else cur_max
    qparams = torch.empty(4, dtype=torch.float32, device=x.device) if qparams is None else qparams
    codes = _codes_like(x, codes_dtype) if quantize else None
    a, im, yo, c, cm, q, ps = dl(x), dl(input_max), dl(yv), dl(codes), dl(cur_max), dl(qparams), dl(per_sample)
    check_call(_lib().fq_forward_online(a.ptr, n_samples, bits, int(bool(signed)), lo_mode, _promo(promotion),
                                        ptr(im), ptr(yo), ptr(c), cm.ptr, q.ptr, ptr(ps), workspace(x.device),
                                        current_stream()))
    return (y, cur_max, qparams, codes) if codes_dtype is not None else (y, cur_max, qparams)


class InputPlan:
    """The input path of ONE converted block as a C-side call plan (fq.h fq_input_plan_*): every argument of
    :func:`forward_online` except the activation's address is captured once; ``run`` then costs an output allocation
    and one foreign call with five integers -- about a third of the host time of building seven DLTensor structs per
    forward (tools/eager_profile.py).  Results are those of :func:`forward_online`: the plan calls the same entry point.

    The plan borrows the block's state tensors; ``matches`` tells whether they (and the activation's shape) are
    still the ones it was built for.
    """
    __slots__ = ("handle", "shape", "device", "dev_index", "cur_max", "qparams", "input_max_ptr", "input_max",
                 "per_sample", "quantize", "sig", "_run", "_keep", "__weakref__")

    def __init__(self, x, bits, signed, lo_mode, input_max=None, quantize=True, cur_max=None, qparams=None,
                 per_sample=None, n_samples=None, promotion=None, sig=None):
        x = _f32(x)
        lib = _lib()
        n_samples = x.shape[0] if n_samples is None else n_samples
        a, im, cm, q, ps = dl(x), dl(input_max), dl(cur_max), dl(qparams), dl(per_sample)
        h = _ffi._c.c_void_p()
        check_call(lib.fq_input_plan_create(a.ptr, n_samples, bits, int(bool(signed)), lo_mode, _promo(promotion),
                                            ptr(im), int(bool(quantize)), cm.ptr, ptr(q), ptr(ps), _ffi._c.byref(h)))
        self.handle = h.value
        self.shape = x.shape
        self.device = x.device
        self.dev_index = x.device.index or 0
        self.cur_max, self.qparams, self.per_sample, self.input_max = cur_max, qparams, per_sample, input_max
        self.input_max_ptr = None if input_max is None else input_max.data_ptr()
        self.quantize = bool(quantize)
        self.sig = sig
        self._run = lib.fq_input_plan_run
        self._keep = (cur_max, qparams, per_sample, input_max)      # the plan holds raw addresses of these

    def __del__(self):
        h, self.handle = self.handle, None
        if h:
            try:
                _lib().fq_input_plan_destroy(h)
            except Exception:       # interpreter shutdown
                pass

    def run(self, x, out=None):
        """x: float32 CUDA tensor of the plan's shape (checked by the caller through ``shape``).  Returns y (None
        for a range-only plan)."""
        # elementwise with whole samples as rows: any dense layout that keeps the samples outermost is read and
        # written as it lies (a channels_last activation gives a channels_last result); anything else is copied
        if not x.is_contiguous():
            if not (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)) or (
                    out is not None and out.stride() != x.stride()):
                x = x.contiguous()
        y = None
        if self.quantize:
            y = torch.empty_like(x) if out is None else out
            if y.stride() != x.stride():
                raise _ffi.FQError("InputPlan.run: out must have the memory layout of x")
        dev = self.dev_index
        raw = _ffi._raw_stream(dev)
        if self._run(self.handle, x.data_ptr(), 0 if y is None else y.data_ptr(), _ffi.workspace_for(dev, raw), raw):
            check_call(-1)
        return y


def forward_from_maxima(x, maxima, bits=8, signed=False, lo_mode=LO_ZERO, out=None, cur_max=None, qparams=None,
                        promotion=None):
    """Online input path with per-sample maxima supplied by the caller (data parallel: the all-gathered maxima of
    the global batch): Kahan mean, scale and quantise in one launch for latency-bound tensors.
    Returns (y, cur_max, qparams)."""
    x = _f32(x)
    out = torch.empty_like(x) if out is None else out
    cur_max = torch.empty(1, dtype=torch.float32, device=x.device) if cur_max is None else cur_max
    qparams = torch.empty(4, dtype=torch.float32, device=x.device) if qparams is None else qparams
    a, m, o, c, q = dl(x), dl(_f32(maxima, "maxima")), dl(out), dl(cur_max), dl(qparams)
    check_call(_lib().fq_forward_from_maxima(a.ptr, m.ptr, bits, int(bool(signed)), lo_mode, _promo(promotion), o.ptr,
                                             None, c.ptr, q.ptr, current_stream()))
    return out, cur_max, qparams


def quant_weight(w, rows, bits, gamma=None, beta=None, mean=None, var=None, bias=None, out=None, bias_out=None,
                 scale_out=None, codes_dtype=None):
    """BN fold (optional) + per-row absmax + scale + quantise: two launches, nothing read back.

    convert_conv2d.py:47-51, 70-95; bits <= 0 folds only (merge_bn.py:65-74).
    Returns (w_q, bias_folded or None, scales or None[, codes]).
    """
    w = _f32(w, "w")
    out = torch.empty_like(w) if out is None else out
    fold = gamma is not None
    if fold and bias_out is None:
        bias_out = torch.empty(w.shape[0], dtype=torch.float32, device=w.device)
    if bits > 0 and scale_out is None:
        scale_out = torch.empty(rows, dtype=torch.float32, device=w.device)
    codes = _codes_like(w, codes_dtype)
    a = dl(w)
    g, b, m, v, bi = dl(gamma), dl(beta), dl(mean), dl(var), dl(bias)
    o, bo, so, c = dl(out), dl(bias_out), dl(scale_out), dl(codes)
    check_call(_lib().fq_quant_weight(a.ptr, rows, bits, ptr(g), ptr(b), ptr(m), ptr(v), ptr(bi), o.ptr, ptr(bo),
                                      ptr(so), ptr(c), workspace(w.device), current_stream()))
    return (out, bias_out, scale_out, codes) if codes_dtype is not None else (out, bias_out, scale_out)


def quant_weight_wino(w, G, GI, GTI, bits, out=None, scale_out=None):
    """Winograd-domain per-channel weight fake-quant: U = G w G^T, per-Cout absmax, quantise, back through the
    pseudo-inverses (convert_conv2d.py:71-83).  Returns (w_q, scales [Cout])."""
    w = _f32(w, "w")
    out = torch.empty_like(w) if out is None else out
    if scale_out is None:
        scale_out = torch.empty(w.shape[0], dtype=torch.float32, device=w.device)
    a, g, gi, gti, o, so = dl(w), dl(G), dl(GI), dl(GTI), dl(out), dl(scale_out)
    check_call(_lib().fq_quant_weight_wino(a.ptr, g.ptr, gi.ptr, gti.ptr, bits, o.ptr, so.ptr, workspace(w.device),
                                           current_stream()))
    return out, scale_out


def wino_backward(dwq, G, GI, GTI, out=None):
    """Straight-through backward of :func:`quant_weight_wino`: dw = G^T (GI^T dwq GTI^T) G."""
    dwq = _f32(dwq, "dwq")
    out = torch.empty_like(dwq) if out is None else out
    a, g, gi, gti, o = dl(dwq), dl(G), dl(GI), dl(GTI), dl(out)
    check_call(_lib().fq_wino_backward(a.ptr, g.ptr, gi.ptr, gti.ptr, o.ptr, current_stream()))
    return out


class WeightPlan:
    """Cached job table for :func:`quant_weight_multi`: the weights (and BN vectors) of a network are persistent
    tensors, so their DLTensor structs are built once; per call only the three flat output buffers change.

    ``jobs``: list of dicts with keys w, rows, bits and optionally gamma, beta, mean, var, bias.
    """

    def __init__(self, jobs):
        self.n = len(jobs)
        self.keep = []
        self.table = (_ffi.FqWeightJob * self.n)()
        self.w_slices, self.bias_slices, self.scale_slices = [], [], []
        w_off = b_off = s_off = 0
        for i, jb in enumerate(jobs):
            w = jb["w"]
            fold = jb.get("gamma") is not None
            rec = self.table[i]
            for name in ("w", "gamma", "beta", "mean", "var", "bias"):
                t = jb.get(name)
                if t is None:
                    setattr(rec, name, None)
                else:
                    t = t.detach()
                    if t.dtype != torch.float32 or not t.is_contiguous():
                        # a .contiguous() copy would be a snapshot that in-place optimizer updates never reach
                        raise _ffi.FQError("WeightPlan: %s of job %d must be contiguous float32 (got %s, strides %s)"
                                           % (name, i, t.dtype, tuple(t.stride())))
                    arg = dl(t)
                    self.keep.append((arg, t))
                    setattr(rec, name, _ffi._c.pointer(arg.t))
            bits = int(jb["bits"])
            rows = int(jb["rows"]) if bits > 0 else 1
            rec.rows, rec.bits = rows, bits
            rec.w_off = w_off
            self.w_slices.append((w_off, w.numel(), tuple(w.shape)))
            w_off += (w.numel() + 3) // 4 * 4
            if fold:
                rec.bias_off = b_off
                self.bias_slices.append((b_off, w.shape[0]))
                b_off += w.shape[0]
            else:
                rec.bias_off = -1
                self.bias_slices.append(None)
            if bits > 0:
                rec.scale_off = s_off
                self.scale_slices.append((s_off, rows))
                s_off += rows
            else:
                rec.scale_off = -1
                self.scale_slices.append(None)
        self.w_total, self.bias_total, self.scale_total = w_off, b_off, s_off
        self.device = jobs[0]["w"].device
        self.ptrs = self.pointers(jobs)

    @staticmethod
    def pointers(jobs):
        """Identity of every tensor a plan holds: the plan is stale as soon as any of them is replaced."""
        return tuple((None if jb.get(name) is None else jb[name].data_ptr())
                     for jb in jobs for name in ("w", "gamma", "beta", "mean", "var", "bias"))

    def valid_for(self, jobs):
        return len(jobs) == self.n and self.pointers(jobs) == self.ptrs


def weight_multi_buffers(plan):
    """(w_flat, b_flat, s_flat, [w_q views], [folded-bias views or None], [scale views or None]) for one call of
    :func:`quant_weight_multi`; a caller that does not keep results across calls may reuse them."""
    dev = plan.device
    w_flat = torch.empty(plan.w_total, dtype=torch.float32, device=dev)
    b_flat = torch.empty(max(plan.bias_total, 1), dtype=torch.float32, device=dev)
    s_flat = torch.empty(max(plan.scale_total, 1), dtype=torch.float32, device=dev)
    ws = [w_flat[o:o + n].view(shape) for o, n, shape in plan.w_slices]
    bs = [None if sl is None else b_flat[sl[0]:sl[0] + sl[1]] for sl in plan.bias_slices]
    ss = [None if sl is None else s_flat[sl[0]:sl[0] + sl[1]] for sl in plan.scale_slices]
    return w_flat, b_flat, s_flat, ws, bs, ss, (dl(w_flat), dl(b_flat), dl(s_flat))


def quant_weight_multi(plan, buffers=None):
    """Run every job of ``plan`` (two launches per 30 jobs).  Returns lists (w_q, bias_folded or None, scales or
    None), views into three flat buffers -- freshly allocated unless ``buffers`` (from :func:`weight_multi_buffers`)
    is given; results equal :func:`quant_weight` per block."""
    if buffers is None:
        buffers = weight_multi_buffers(plan)
    wf, bf, sf = buffers[6]
    dev = plan.device.index or 0
    check_call(_lib().fq_quant_weight_multi(plan.table, plan.n, wf.ptr, bf.ptr, sf.ptr, workspace(dev), current_stream()))
    return buffers[3], buffers[4], buffers[5]


def fold_backward_multi(jobs):
    """Backward of the fake-BN fold of many blocks in ONE launch.  ``jobs``: dicts with dwq, w, gamma, mean, var and
    optionally dbq, bias (all float32, contiguous).  Returns per job (dw, dgamma, dbias or None, dbeta or None); the
    outputs are views of three freshly allocated flat buffers."""
    dev = jobs[0]["w"].device
    n_w = sum(jb["w"].numel() for jb in jobs)
    n_c = sum(jb["w"].shape[0] for jb in jobs)
    dw_flat = torch.empty(n_w, dtype=torch.float32, device=dev)
    vec_flat = torch.empty(3 * n_c, dtype=torch.float32, device=dev)
    table = (_ffi.FqFoldBwdJob * len(jobs))()
    keep, outs = [], []
    wo = co = 0
    for rec, jb in zip(table, jobs):
        w = jb["w"]
        c, n = w.shape[0], w.numel()
        has_b = jb.get("dbq") is not None
        dw = dw_flat[wo:wo + n].view(w.shape)
        dgamma, dbias, dbeta = vec_flat[co:co + c], vec_flat[n_c + co:n_c + co + c], vec_flat[2 * n_c + co:2 * n_c + co + c]
        wo += n
        co += c
        fields = dict(dwq=_f32(jb["dwq"], "dwq"), dbq=jb.get("dbq"), w=w.detach(), gamma=jb["gamma"].detach(),
                      mean=jb["mean"].detach(), var=jb["var"].detach(),
                      bias=None if jb.get("bias") is None else jb["bias"].detach(),
                      dw=dw, dgamma=dgamma, dbias=dbias if has_b else None, dbeta=dbeta if has_b else None)
        for name, t in fields.items():
            if t is None:
                setattr(rec, name, None)
            else:
                arg = dl(_f32(t, name))
                keep.append(arg)
                setattr(rec, name, _ffi._c.pointer(arg.t))
        outs.append((dw, dgamma, dbias if has_b else None, dbeta if has_b else None))
    check_call(_lib().fq_fold_backward_multi(table, len(jobs), current_stream()))
    return outs


class FoldBackwardPlan:
    """Cached job table for the fold backward of a fixed set of fake-BN blocks: the persistent tensors (w, gamma, mean,
    var, bias) are described once; per call only the ADDRESSES of the incoming gradients and of the freshly allocated
    outputs are patched into DLTensor structs that already exist.  Same entry point and results as
    :func:`fold_backward_multi`; about a tenth of its host time for MobileNetV2's 52 blocks (tools/eager_profile_qat.py).

    ``jobs``: dicts with w, gamma, mean, var and optionally bias (contiguous float32 CUDA tensors)."""

    def __init__(self, jobs):
        c = _ffi._c
        self.n = len(jobs)
        self.device = jobs[0]["w"].device
        dev = self.device.index or 0
        self.table = (_ffi.FqFoldBwdJob * self.n)()
        self.keep = []
        self.dyn = []                   # per job: {field: (DLTensor, pointer)} of the tensors whose address changes
        self.w_shapes, self.w_sizes, self.c_sizes = [], [], []
        for rec, jb in zip(self.table, jobs):
            w = jb["w"]
            for name in ("w", "gamma", "mean", "var", "bias"):
                t = jb.get(name)
                if t is None:
                    setattr(rec, name, None)
                    continue
                t = t.detach()
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise _ffi.FQError("FoldBackwardPlan: %s must be contiguous float32" % name)
                arg = dl(t)
                self.keep.append((arg, t))
                setattr(rec, name, c.pointer(arg.t))
            cout = w.shape[0]
            d = {}
            for name, shape in (("dwq", tuple(w.shape)), ("dw", tuple(w.shape)), ("dbq", (cout,)), ("dgamma", (cout,)),
                                ("dbias", (cout,)), ("dbeta", (cout,))):
                sh = (c.c_int64 * len(shape))(*shape)
                t = _ffi.DLTensor(None, _ffi.DLDevice(_ffi.kDLCUDA, dev), len(shape), _ffi.DLDataType(2, 32, 1), sh, None, 0)
                ptr_ = c.pointer(t)
                self.keep.append((sh, t, ptr_))
                d[name] = (t, ptr_)
            for name in ("dwq", "dw", "dgamma"):
                setattr(rec, name, d[name][1])
            self.dyn.append(d)
            self.w_shapes.append(tuple(w.shape))
            self.w_sizes.append(w.numel())
            self.c_sizes.append(cout)
        self.n_w, self.n_c = sum(self.w_sizes), sum(self.c_sizes)

    def run(self, dwqs, dbqs):
        """dwqs[i]: gradient w.r.t. the quantised folded weight of job i; dbqs[i]: w.r.t. its folded bias, or None.
        Returns per job (dw, dgamma, dbias or None, dbeta or None)."""
        dev = self.device
        dw_flat = torch.empty(self.n_w, dtype=torch.float32, device=dev)
        dg_flat = torch.empty(self.n_c, dtype=torch.float32, device=dev)
        db_flat = torch.empty(self.n_c, dtype=torch.float32, device=dev)
        dbe_flat = torch.empty(self.n_c, dtype=torch.float32, device=dev)
        dws = dw_flat.split_with_sizes(self.w_sizes)
        dgs, dbs, dbes = (f.split_with_sizes(self.c_sizes) for f in (dg_flat, db_flat, dbe_flat))
        p_w, p_g, p_b, p_be = dw_flat.data_ptr(), dg_flat.data_ptr(), db_flat.data_ptr(), dbe_flat.data_ptr()
        wo = co = 0
        keep, outs = [], []
        for i, (rec, d) in enumerate(zip(self.table, self.dyn)):
            dwq = dwqs[i]
            if dwq.dtype is not torch.float32 or not dwq.is_contiguous():
                dwq = _f32(dwq, "dwq")
                keep.append(dwq)
            d["dwq"][0].data = dwq.data_ptr()
            d["dw"][0].data = p_w + 4 * wo
            d["dgamma"][0].data = p_g + 4 * co
            dbq = dbqs[i]
            if dbq is None:
                rec.dbq = rec.dbias = rec.dbeta = None
                outs.append((dws[i].view(self.w_shapes[i]), dgs[i], None, None))
            else:
                if dbq.dtype is not torch.float32 or not dbq.is_contiguous():
                    dbq = _f32(dbq, "dbq")
                    keep.append(dbq)
                d["dbq"][0].data = dbq.data_ptr()
                d["dbias"][0].data = p_b + 4 * co
                d["dbeta"][0].data = p_be + 4 * co
                rec.dbq, rec.dbias, rec.dbeta = d["dbq"][1], d["dbias"][1], d["dbeta"][1]
                outs.append((dws[i].view(self.w_shapes[i]), dgs[i], dbs[i], dbes[i]))
            wo += self.w_sizes[i]
            co += self.c_sizes[i]
        check_call(_lib().fq_fold_backward_multi(self.table, self.n, current_stream()))
        return outs


# ---- K3 ------------------------------------------------------------------------------------------
def ste_backward(dy, x=None, qparams=None, mode=STE_IDENTITY):
    """ste_func.py:43-44.  Identity aliases dy (zero bytes moved); the clip mask is an extension."""
    if mode == STE_IDENTITY:
        return dy
    dy = _f32(dy, "dy")
    dx = torch.empty_like(dy)
    g, a, q, o = dl(dy), dl(_f32(x)), dl(qparams), dl(dx)
    check_call(_lib().fq_ste_backward(g.ptr, a.ptr, q.ptr, o.ptr, mode, current_stream()))
    return dx


# ---- K4 ------------------------------------------------------------------------------------------
def ema_update(state, cur, momentum=0.9, scalar_cur=True, promotion=None):
    """state <- (1 - m) * cur + m * state, in place.  convert.py:66-78."""
    s, c = dl(state), dl(_f32(cur, "cur"))
    check_call(_lib().fq_ema_update(s.ptr, c.ptr, float(momentum), int(bool(scalar_cur)), _promo(promotion),
                                    current_stream()))
    return state


# ---- K5 ------------------------------------------------------------------------------------------
def hist_nonzero(x, max_, bins, counts, promotion=None, bad_flag=None):
    """counts[bin] += 1 over the clipped non-zero elements.  distribution_calibrate.py:39-45.
    ``bad_flag`` (int32 [1]) is raised when the reference's asserts (:35-36) would fail on this tensor."""
    a, m, c, f = dl(_act(x)), dl(max_), dl(counts), dl(bad_flag)
    check_call(_lib().fq_hist_nonzero(a.ptr, m.ptr, bins, _promo(promotion), c.ptr, ptr(f), current_stream()))
    return counts


def hist_nonzero_multi(xs, maxes, max_stride, max_offset, bins, counts, promotion=None, bad_flags=None):
    """Histograms of several layer inputs in one launch; ``counts`` is int64 [len(xs), bins + 1] and the
    frozen max of ``xs[i]`` is ``maxes.view(-1)[i * max_stride + max_offset]``; ``bad_flags``: int32 [len(xs)]."""
    args = [dl(_act(x)) for x in xs]
    arr = (_ffi.P * len(args))(*[_ffi._c.pointer(a.t) for a in args])
    m, c, f = dl(maxes), dl(counts), dl(bad_flags)
    check_call(_lib().fq_hist_nonzero_multi(arr, len(args), m.ptr, max_stride, max_offset, bins, _promo(promotion),
                                            c.ptr, ptr(f), current_stream()))
    return counts


def hist_accumulate(counts, hist, first, seen_last=None):
    """hist (+)= float32(counts); counts <- 0.  distribution_calibrate.py:47,103-104."""
    c, h, s = dl(counts), dl(hist), dl(seen_last)
    check_call(_lib().fq_hist_accumulate_f32(c.ptr, h.ptr, int(bool(first)), ptr(s), current_stream()))
    return hist


def kl_search(hist, levels, min_bins, bins, promotion=None, divergence=None, margin=None):
    """Best threshold bin per histogram (hist: [n_data] or [layers, n_data]).  :117-171.

    Returns (best int32 [layers], divergence float64 [layers, bins]); ``margin`` (float64 [layers]) receives the
    relative gap between the best and the runner-up divergence (see :data:`KL_TIE_MARGIN`)."""
    hist = _f32(hist, "hist")
    layers = 1 if hist.dim() == 1 else hist.shape[0]
    best = torch.empty(layers, dtype=torch.int32, device=hist.device)
    if divergence is None:
        divergence = torch.full((layers, bins), float("nan"), dtype=torch.float64, device=hist.device)
    h, b, d, m = dl(hist), dl(best), dl(divergence), dl(margin)
    check_call(_lib().fq_kl_search(h.ptr, levels, min_bins, bins, _promo(promotion), b.ptr, d.ptr, ptr(m),
                                   current_stream()))
    return best, divergence


# A chosen bin whose divergence beats the runner-up by less than this (relative) is reported as a near-tie: the
# divergences are float64 sums of p*log(p/q) and agree with NumPy's to ~1e-13 relative, so above the margin the
# arg-min is the reference's; below it the reference's own answer hinges on the last place of its libm's log.
KL_TIE_MARGIN = 1e-9


def kl_threshold(best, fm_max, bins, out=None):
    """input_max = (best + 0.5) * (fm_max / bins).  simulate_quantization.py:310."""
    out = torch.empty(best.numel(), dtype=torch.float32, device=best.device) if out is None else out
    b, m, o = dl(best), dl(_f32(fm_max, "fm_max")), dl(out)
    check_call(_lib().fq_kl_threshold(b.ptr, m.ptr, bins, o.ptr, current_stream()))
    return out


# ---- K6 ------------------------------------------------------------------------------------------
def quantize_int8_export(w, range2):
    """MXNet contrib.quantize(out_type='int8').  freeze.py:100-103 -> (int8, {-real, +real})."""
    w = _f32(w, "w")
    out = torch.empty(w.shape, dtype=torch.int8, device=w.device)
    out_range = torch.empty(2, dtype=torch.float32, device=w.device)
    a, r, o, orr = dl(w), dl(_f32(range2, "range2")), dl(out), dl(out_range)
    check_call(_lib().fq_quantize_int8_export(a.ptr, r.ptr, o.ptr, orr.ptr, current_stream()))
    return out, out_range


def qconv_quantize(x, range2):
    """nn/quantized_conv.py:54-61 -> (int32 codes, scale (1,))."""
    x = _f32(x)
    codes = torch.empty(x.shape, dtype=torch.int32, device=x.device)
    scale = torch.empty(1, dtype=torch.float32, device=x.device)
    a, r, c, s = dl(x), dl(_f32(range2, "range2")), dl(codes), dl(scale)
    check_call(_lib().fq_qconv_quantize(a.ptr, r.ptr, c.ptr, s.ptr, current_stream()))
    return codes, scale


def qconv_dequantize(acc, s_in, s_w):
    """nn/quantized_conv.py:74-76."""
    acc = acc.contiguous()
    y = torch.empty(acc.shape, dtype=torch.float32, device=acc.device)
    a, si, sw, o = dl(acc), dl(s_in), dl(s_w), dl(y)
    check_call(_lib().fq_qconv_dequantize(a.ptr, si.ptr, sw.ptr, o.ptr, current_stream()))
    return y


# ---- QConv2D on the tensor cores ---------------------------------------------------------------------
def qconv_pack_input(x, range2, pad_h, pad_w, unsigned=False):
    """fp32 NCHW -> spatially padded NHWC 8-bit codes + scale (1,).  nn/quantized_conv.py:108-116, 54-61."""
    x = _f32(x)
    n, c, h, w = x.shape
    xq = torch.empty((n, h + 2 * pad_h, w + 2 * pad_w, c), dtype=torch.uint8 if unsigned else torch.int8, device=x.device)
    scale = torch.empty(1, dtype=torch.float32, device=x.device)
    a, r, q, s = dl(x), dl(_f32(range2, "range2")), dl(xq), dl(scale)
    check_call(_lib().fq_qconv_pack_input(a.ptr, r.ptr, int(pad_h), int(pad_w), q.ptr, s.ptr, current_stream()))
    return xq, scale


def qconv_pack_weight(w, range2):
    """fp32 [Cout, Cg, KH, KW] -> int8 codes [Cout, KH, KW, Cg] + scale (1,)."""
    w = _f32(w, "w")
    co, cg, kh, kw = w.shape
    wq = torch.empty((co, kh, kw, cg), dtype=torch.int8, device=w.device)
    scale = torch.empty(1, dtype=torch.float32, device=w.device)
    a, r, q, s = dl(w), dl(_f32(range2, "range2")), dl(wq), dl(scale)
    check_call(_lib().fq_qconv_pack_weight(a.ptr, r.ptr, q.ptr, s.ptr, current_stream()))
    return wq, scale


def qconv_igemm(xq, wq, bias_q, s_in, s_w, strides, groups=1, relu=False):
    """Integer convolution on tcgen05 with the bias / ReLU / dequantise epilogue fused.  nn/quantized_conv.py:143-158.
    ``bias_q``: int32 codes, or the float32 bias itself (quantised in the epilogue with b_scale = s_in * s_w, :122-127)."""
    n, hp, wp, _ = xq.shape
    co, kh, kw, _ = wq.shape
    ho, wo = (hp - kh) // strides[0] + 1, (wp - kw) // strides[1] + 1
    out = torch.empty((n, co, ho, wo), dtype=torch.float32, device=xq.device)
    a, b, c, si, sw, o = dl(xq), dl(wq), dl(bias_q), dl(s_in), dl(s_w), dl(out)
    check_call(_lib().fq_qconv_igemm(a.ptr, b.ptr, ptr(c), si.ptr, sw.ptr, int(strides[0]), int(strides[1]), int(groups),
                                     int(bool(relu)), o.ptr, current_stream()))
    return out
