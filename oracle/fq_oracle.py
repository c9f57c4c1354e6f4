"""CPU oracle for the fake-quantization hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a dtype-explicit NumPy restatement of the arithmetic that
hey-yahei/Quantization.MXNet executes on its CPU path.  It is the *checker*
for the CUDA kernels: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product package (``quantization.mxnet_b200``) never does.

Parity status
-------------
* ``discrete_histogram`` / ``accumulate_histograms`` / ``kl_calibrate`` are
  PINNED: ``oracle/make_golden.py`` executes the reference's own
  ``quantize/distribution_calibrate.py`` (loaded by file path, NumPy 2.x = NEP 50
  scalar promotion) and commits its outputs under ``tests/golden/``;
  ``tests/test_oracle_golden.py`` checks this restatement against them bit for
  bit.
* Everything that the reference delegates to Apache MXNet 1.x NDArray ops
  (``ste_func.py``, ``convert_conv2d.py``, ``convert_dense.py``, ``convert.py``
  EMA, ``merge_bn.py``, ``nn/quantized_conv.py``, ``freeze.py`` contrib.quantize)
  is **PARITY UNPINNED**: MXNet (unpinned 1.x, >= 1.5) is neither vendored in
  the reference nor installable offline, and the reference's tests hold no
  golden vectors.  The restatement below therefore *defines* the contract;
  every MXNet op semantic it encodes is spelled out next to the function.

Scalar promotion regimes
------------------------
Host-side scalar math in the reference (``max_ / 255``, ``scale + 1e-10``,
``(1 - m) * cur``, ``bins / (max_ + 1e-5)``, builtin ``sum`` start value) mixes
``numpy.float32`` scalars with Python numbers.  NumPy 1.x ("legacy", what MXNet
1.x requires) promotes those to float64; NumPy >= 2 ("nep50", this container)
keeps float32.  Every function that depends on it takes
``promotion="legacy"|"nep50"``.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
F64 = np.float64
EPS_DIV = 1e-10      # ste_func.py:39,41
EPS_BN = 1e-10       # convert_conv2d.py:50-51, merge_bn.py:65-66,74
EPS_HIST = 1e-5      # distribution_calibrate.py:41
_PROMOTIONS = ("legacy", "nep50")


def _check_promotion(promotion):
    if promotion not in _PROMOTIONS:
        raise ValueError("promotion must be 'legacy' or 'nep50', got %r" % (promotion,))


# ---------------------------------------------------------------------------
# MXNet elementwise op semantics (explicit oracle decisions, see module header)
# ---------------------------------------------------------------------------
def roundf(x):
    """mshadow_op::round == C roundf: half away from zero (NOT np.round)."""
    x = np.asarray(x, dtype=F32)
    t = np.trunc(x)
    with np.errstate(invalid="ignore"):
        frac = np.abs(x - t)            # exact in fp32
        up = frac >= F32(0.5)
    return np.where(up, t + np.copysign(F32(1), x), t).astype(F32)


def clip(x, lo, hi):
    """mshadow_op::clip: x>hi -> hi ; x<lo -> lo ; else x (NaN passes)."""
    x = np.asarray(x, dtype=F32)
    lo = F32(lo)
    hi = F32(hi)
    return np.where(x > hi, hi, np.where(x < lo, lo, x)).astype(F32)


def kahan_sum_f32(v):
    """mshadow::red::sum::Reduce with residual, sequential in index order."""
    s = F32(0)
    c = F32(0)
    for a in np.asarray(v, dtype=F32).reshape(-1):
        y = F32(a - c)
        t = F32(s + y)
        c = F32(F32(t - s) - y)
        s = t
    return s


def mean_kahan_f32(v):
    """MXNet CPU ``mean``: Kahan fp32 sum, then one fp32 division by N."""
    v = np.asarray(v, dtype=F32).reshape(-1)
    return F32(kahan_sum_f32(v) / F32(v.size))


# ---------------------------------------------------------------------------
# K1  range reductions
# ---------------------------------------------------------------------------
def absmax_rows(x, rows):
    """max |x| over each of ``rows`` equal contiguous rows.

    convert_conv2d.py:56 (rows=N), :75 (rows=Cout), :86 (rows=G), :92 (rows=1);
    convert_dense.py:41,54,60; quantized_conv.py:65.
    """
    x = np.asarray(x, dtype=F32)
    if x.size == 0:
        return np.zeros((rows,), dtype=F32)
    return np.abs(x.reshape(rows, -1)).max(axis=1).astype(F32)


def minmax(x):
    """quantized_conv.py:68-69 ; distribution_calibrate.py:34-35."""
    x = np.asarray(x, dtype=F32)
    return F32(x.min()), F32(x.max())


def input_range(x):
    """current_input_max = mean_n(max_chw |x|)   convert_conv2d.py:56, convert_dense.py:41."""
    x = np.asarray(x, dtype=F32)
    per_sample = absmax_rows(x, x.shape[0])
    return mean_kahan_f32(per_sample), per_sample


# ---------------------------------------------------------------------------
# scale computation (host scalar math in the reference)
# ---------------------------------------------------------------------------
def qmax_of(bits, signed):
    return (2 ** (bits - 1) - 1) if signed else (2 ** bits - 1)


def input_qparams(max_, bits, signed, promotion="legacy", layer="conv"):
    """(d, s, lo, hi) for the input path.

    convert_conv2d.py:57-66 / convert_dense.py:42-49 + ste_func.py:34,41.
    ``layer="dense"`` reproduces the Dense quirk: no clip_min is passed, so
    clip_min defaults to 0 even for signed inputs (convert_dense.py:49).
    """
    _check_promotion(promotion)
    max_ = F32(max_)
    q = qmax_of(bits, signed)
    if promotion == "legacy":
        s64 = F64(max_) / F64(q)
        d = F32(s64 + EPS_DIV)
        s = F32(s64)
    else:
        s32 = F32(max_ / F32(q))
        d = F32(s32 + F32(EPS_DIV))
        s = s32
    lo = F32(-max_) if (signed and layer == "conv") else F32(0.0)
    return d, s, lo, max_


def weight_scales(w, rows, bits):
    """Per-row (s_r, d_r), all fp32 NDArray math (convert_conv2d.py:70-95)."""
    m = absmax_rows(w, rows)
    s = (m / F32(qmax_of(bits, True))).astype(F32)
    d = (s + F32(EPS_DIV)).astype(F32)
    return s, d, m


# ---------------------------------------------------------------------------
# K2  fake-quant forward
# ---------------------------------------------------------------------------
def fake_quant_scalar(x, d, s, lo=None, hi=None):
    """ste_func.py:41 (clip) / :39 (no clip) with a scalar divisor/multiplier.

    Returns (y, code); ``code`` is the integer-valued fp32 rounded quotient.
    """
    x = np.asarray(x, dtype=F32)
    if hi is not None:
        x = clip(x, lo, hi)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        code = roundf((x / F32(d)).astype(F32))
        y = (code * F32(s)).astype(F32)
    return y, code


def fake_quant_rows(x, rows, s, d=None):
    """ste_func.py:39 with a broadcast per-row scale tensor (weights)."""
    x = np.asarray(x, dtype=F32)
    s = np.asarray(s, dtype=F32).reshape(rows, 1)
    d = (s + F32(EPS_DIV)).astype(F32) if d is None else np.asarray(d, dtype=F32).reshape(rows, 1)
    x2 = x.reshape(rows, -1)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        code = roundf((x2 / d).astype(F32))
        y = (code * s).astype(F32)
    return y.reshape(x.shape), code.reshape(x.shape)


def fake_quant_input(x, bits=8, signed=False, input_max=None, promotion="legacy", layer="conv"):
    """Whole input path of one converted block (online when input_max is None)."""
    cur, _ = input_range(x)
    max_ = cur if input_max is None else F32(input_max)
    d, s, lo, hi = input_qparams(max_, bits, signed, promotion, layer)
    y, code = fake_quant_scalar(x, d, s, lo, hi)
    return y, code, cur, (d, s, lo, hi)


def fake_quant_weight(w, bits=8, quant_type="layer", groups=1):
    """Weight path of convert_conv2d.py:70-95 / convert_dense.py:52-63."""
    w = np.asarray(w, dtype=F32)
    if quant_type == "channel":
        rows = w.shape[0]
    elif quant_type == "group":
        rows = groups
        if rows not in (1, w.shape[0]):
            # (G,1,1,1) does not broadcast against (Cout,...) in MXNet
            raise ValueError("reference broadcast rule: per-group needs G in {1, Cout}")
    else:
        rows = 1
    s, d, m = weight_scales(w, rows, bits)
    y, code = fake_quant_rows(w, rows, s, d)
    return y, code, s


# ---------------------------------------------------------------------------
# Winograd-domain weight quantisation   convert_conv2d.py:71-83 ; wino_matrix.py:28-60
# ---------------------------------------------------------------------------
_WINO_G = {
    "F23": [[1, 0, 0], [1 / 2, 1 / 2, 1 / 2], [1 / 2, -1 / 2, 1 / 2], [0, 0, 1]],
    "F43": [[1 / 4, 0, 0], [-1 / 6, -1 / 6, -1 / 6], [-1 / 6, 1 / 6, -1 / 6], [1 / 24, 1 / 12, 1 / 6],
            [1 / 24, -1 / 12, 1 / 6], [0, 0, 1]],
    "F63": [[1, 0, 0], [-2 / 9, -2 / 9, -2 / 9], [-2 / 9, 2 / 9, -2 / 9], [1 / 90, 1 / 45, 2 / 45],
            [1 / 90, -1 / 45, 2 / 45], [32 / 45, 16 / 45, 8 / 45], [32 / 45, -16 / 45, 8 / 45], [0, 0, 1]],
}


def winograd_matrices(name):
    """G as ``nd.array`` makes it (float32), and the float32 pseudo-inverses of G and G.T that the reference
    computes with ``np.linalg.pinv(G.asnumpy())`` on every forward (convert_conv2d.py:80-82)."""
    G = np.array(_WINO_G[name], dtype=F32)
    GI = np.linalg.pinv(G).astype(F32)
    GTI = np.linalg.pinv(np.ascontiguousarray(G.T)).astype(F32)
    return G, GI, GTI


def _fma32(a, b, c):
    """float32 fused multiply-add.  The product of two float32 is exact in x87 extended precision (64-bit
    significand); the sum is rounded once there and once more to float32 -- a double rounding that differs from a
    true FMA with probability ~2^-39 per operation, far below what the tests can hit."""
    LD = np.longdouble
    assert np.finfo(LD).nmant >= 63, "needs x87 extended precision"
    return (np.asarray(a, F32).astype(LD) * np.asarray(b, F32).astype(LD) + np.asarray(c, F32).astype(LD)).astype(F32)


def _dot_chain(pairs):
    """sum_k a_k * b_k in ascending k as ``acc = a0*b0; acc = fma(a_k, b_k, acc)``.

    PARITY UNPINNED: MXNet's ``nd.dot`` is a BLAS sgemm (contraction length 3 / 4 / 6 / 8 here) whose summation
    order and FMA use are not defined by MXNet; this chain is the contract the CUDA kernel is held to."""
    (a0, b0), rest = pairs[0], pairs[1:]
    acc = (np.asarray(a0, F32) * np.asarray(b0, F32)).astype(F32)
    for a, b in rest:
        acc = _fma32(a, b, acc)
    return acc


def wino_transform(w, G):
    """U = (G w) G^T per 3x3 kernel (:72-73); w: [..., 3, 3] -> [..., a, a]."""
    w = np.asarray(w, dtype=F32)
    a = G.shape[0]
    # (G w)[p][c] = sum_r G[p][r] w[r][c]
    T = _dot_chain([(G[:, r].reshape(a, 1), w[..., r, :][..., None, :]) for r in range(3)])          # [..., a, 3]
    # U[p][q] = sum_c (G w)[p][c] G[q][c]
    return _dot_chain([(T[..., :, c][..., :, None], G[:, c].reshape(1, a)) for c in range(3)])       # [..., a, a]


def fake_quant_weight_wino(w, name, bits=8):
    """convert_conv2d.py:71-83 for quant_type='channel' and a 3x3 kernel -> (w_q, scales [Cout], U, Uq)."""
    w = np.asarray(w, dtype=F32)
    assert w.ndim == 4 and w.shape[2:] == (3, 3)
    G, GI, GTI = winograd_matrices(name)
    a = G.shape[0]
    U = wino_transform(w, G)                                    # [Cout, Cin, a, a]
    s, d, _ = weight_scales(U, w.shape[0], bits)
    Uq, _ = fake_quant_rows(U, w.shape[0], s, d)
    # (G+ Uq)[r][q] = sum_p GI[r][p] Uq[p][q]
    V = _dot_chain([(GI[:, p].reshape(3, 1), Uq[..., p, :][..., None, :]) for p in range(a)])        # [..., 3, a]
    # w_q[r][c] = sum_q (G+ Uq)[r][q] GTI[q][c]
    wq = _dot_chain([(V[..., :, q][..., :, None], GTI[q, :].reshape(1, 3)) for q in range(a)])       # [..., 3, 3]
    return wq, s, U, Uq


def wino_backward(dwq, name):
    """Straight-through backward of the above in the order autograd replays it:
    dX = dw_q GTI^T ; dUq = GI^T dX ; dT = dU G ; dw = G^T dT."""
    dwq = np.asarray(dwq, dtype=F32)
    G, GI, GTI = winograd_matrices(name)
    a = G.shape[0]
    dX = _dot_chain([(dwq[..., :, c][..., :, None], GTI[:, c].reshape(1, a)) for c in range(3)])     # [..., 3, a]
    dU = _dot_chain([(GI[r, :].reshape(a, 1), dX[..., r, :][..., None, :]) for r in range(3)])       # [..., a, a]
    dT = _dot_chain([(dU[..., :, q][..., :, None], G[q, :].reshape(1, 3)) for q in range(a)])        # [..., a, 3]
    return _dot_chain([(G[p, :].reshape(3, 1), dT[..., p, :][..., None, :]) for p in range(a)])      # [..., 3, 3]


# ---------------------------------------------------------------------------
# BN fold   convert_conv2d.py:47-51 ; merge_bn.py:65-74
# ---------------------------------------------------------------------------
def fold_bn(w, bias, gamma, beta, mean, var):
    w = np.asarray(w, dtype=F32)
    cout = w.shape[0]
    gamma = np.asarray(gamma, dtype=F32)
    beta = np.asarray(beta, dtype=F32)
    mean = np.asarray(mean, dtype=F32)
    var = np.asarray(var, dtype=F32)
    bias = np.zeros(cout, dtype=F32) if bias is None else np.asarray(bias, dtype=F32)
    sd = np.sqrt((var + F32(EPS_BN)).astype(F32)).astype(F32)
    w2 = ((w.reshape(cout, -1) * gamma.reshape(-1, 1)).astype(F32) / sd.reshape(-1, 1)).astype(F32)
    b2 = (((gamma * (bias - mean).astype(F32)).astype(F32) / sd).astype(F32) + beta).astype(F32)
    return w2.reshape(w.shape), b2


# ---------------------------------------------------------------------------
# fake-BN batch statistics   convert_conv2d.py:144-154
# ---------------------------------------------------------------------------
def kahan_sum_rows_f32(a):
    """Kahan fp32 sum of every row of a [rows, n] array, sequential in column order (vectorised over the rows):
    MXNet's CPU ``sum(axis=...)`` walks the reduced coordinates in row-major order with mshadow::red::sum's
    residual (broadcast_reduce-inl.h seq_reduce_compute)."""
    a = np.asarray(a, dtype=F32)
    s = np.zeros(a.shape[0], F32)
    c = np.zeros(a.shape[0], F32)
    for k in range(a.shape[1]):
        y = (a[:, k] - c).astype(F32)
        t = (s + y).astype(F32)
        c = ((t - s).astype(F32) - y).astype(F32)
        s = t
    return s


def channel_stats(y):
    """convert_conv2d.py:150-153 on the raw conv output y [N, C, H, W] -> (current_mean, current_var) fp32 [C]:
        num = N*H*W ; mean = y.sum(axis=(0,2,3)) / num ; var = ((y - mean) ** 2).sum(axis=(0,2,3)) / num
    PARITY UNPINNED (MXNet ops): sums = sequential Kahan fp32 over (n, h, w); ``/ num`` = _div_scalar by fl32(num);
    ``** 2`` = _power_scalar, taken as the fp32 product a*a."""
    y = np.asarray(y, dtype=F32)
    n, c = y.shape[0], y.shape[1]
    flat = np.moveaxis(y.reshape(n, c, -1), 1, 0).reshape(c, -1)           # [C, N*H*W] in (n, h, w) order
    num = F32(flat.shape[1])
    mean = (kahan_sum_rows_f32(flat) / num).astype(F32)
    diff = (flat - mean[:, None]).astype(F32)
    var = (kahan_sum_rows_f32((diff * diff).astype(F32)) / num).astype(F32)
    return mean, var


# ---------------------------------------------------------------------------
# K3  STE backward   ste_func.py:43-44
# ---------------------------------------------------------------------------
def ste_backward(dy, x=None, lo=None, hi=None, mode="identity"):
    dy = np.asarray(dy, dtype=F32)
    if mode == "identity":
        return dy
    x = np.asarray(x, dtype=F32)
    return np.where((x >= F32(lo)) & (x <= F32(hi)), dy, F32(0)).astype(F32)


# ---------------------------------------------------------------------------
# K4  EMA   convert.py:66-78
# ---------------------------------------------------------------------------
def ema_scalar(state, cur, momentum=0.9, promotion="legacy"):
    """input_max <- (1-m)*cur + m*input_max with cur a host numpy.float32."""
    _check_promotion(promotion)
    a = 1 - momentum                       # Python double
    state = np.asarray(state, dtype=F32)
    cur = np.asarray(cur, dtype=F32)
    if promotion == "legacy":
        t = (a * cur.astype(F64)).astype(F32)
    else:
        t = (F32(a) * cur).astype(F32)
    return ((F32(momentum) * state).astype(F32) + t).astype(F32)


def ema_tensor(state, cur, momentum=0.9):
    """running_mean/var <- (1-m)*cur + m*state with cur an NDArray (all fp32)."""
    state = np.asarray(state, dtype=F32)
    cur = np.asarray(cur, dtype=F32)
    return ((F32(1 - momentum) * cur).astype(F32) + (F32(momentum) * state).astype(F32)).astype(F32)


# ---------------------------------------------------------------------------
# K5  histogram + KL search   distribution_calibrate.py
# ---------------------------------------------------------------------------
def hist_scale(max_, bins, promotion="nep50"):
    _check_promotion(promotion)
    max_ = F32(max_)
    if promotion == "legacy":
        return F32(F64(bins) / (F64(max_) + EPS_HIST))
    return F32(F32(bins) / F32(max_ + F32(EPS_HIST)))


def histogram_counts(x, bins, max_, promotion="nep50"):
    """Integer counts of one batch; length bins or bins+1 (2049th-bin quirk).

    distribution_calibrate.py:39-45: clip to [0,max_], drop zeros, multiply by
    ``bins/(max_+1e-5)``, truncate to int32, bincount(minlength=bins).
    """
    v = np.asarray(x, dtype=F32).reshape(-1)
    max_ = F32(max_)
    v = np.minimum(np.maximum(v, F32(0)), max_)   # ndarray.clip(0, max_)
    v = v[v != 0]
    sc = hist_scale(max_, bins, promotion)
    q = (v * sc).astype(F32).astype(np.int32)
    return np.bincount(q, minlength=bins).astype(np.int64)


def discrete_histogram(x, bins, max_=None, promotion="nep50"):
    """distribution_calibrate.py:31-47 -> (hist float32, max_)."""
    x = np.asarray(x, dtype=F32)
    if max_ is None:
        max_ = F32(x.max())
    if not (x.min() >= 0):
        raise AssertionError("Activation should >=0")
    if not (max_ > 0):
        raise AssertionError("Bad distribution: all zero-value")
    return histogram_counts(x, bins, max_, promotion).astype(F32), F32(max_)


def accumulate_histograms(batches, bins, promotion="nep50"):
    """collect_feature_maps' per-block bookkeeping (:95-104): first batch's max
    is frozen, per-batch float32(count) is added in fp32 in batch order."""
    hist = 0
    fm_max = None
    for fm in batches:
        h, m = discrete_histogram(fm, bins, fm_max, promotion)
        if fm_max is None:
            fm_max = m
        hist = hist + h            # raises ValueError if lengths differ, as NumPy does
    return hist, fm_max


def _seq_sum(a, dtype):
    """Python builtin ``sum`` (strictly left-to-right) with the given accumulator."""
    a = np.asarray(a)
    if a.size == 0:
        return dtype(0)
    return np.cumsum(a, dtype=dtype)[-1]


def kl_divergences(data, levels, min_bins, bins, promotion="nep50"):
    """Divergence of every candidate i in [min_bins, bins) (float64, may be NaN).

    Follows distribution_calibrate.py:138-166 operation by operation; the
    sequential Python ``sum`` calls become ``cumsum`` with the accumulator dtype
    the regime implies (float32 under nep50, float64 under legacy for the two
    sums over the float32 histogram).
    """
    _check_promotion(promotion)
    data = np.asarray(data, dtype=F32)
    acc = F32 if promotion == "nep50" else F64
    out = np.full((bins,), np.nan, dtype=F64)
    with np.errstate(all="ignore"):
        for i in range(min_bins, bins):
            ref = data[:i].copy()
            tail = _seq_sum(data[i:], acc)
            ref[i - 1] = F32(acc(ref[i - 1]) + tail)
            total = _seq_sum(ref, acc)
            ref = (ref / F32(total)).astype(F32)          # both regimes divide in fp32
            t = (np.arange(i, dtype=np.int64) * levels) / F64(i)     # float64
            fl = t.astype(np.int32)
            cand = np.zeros(levels, dtype=F64)
            np.add.at(cand, fl, data[:i].astype(F64))
            ce = np.clip(np.ceil(t), 0, levels - 1).astype(np.int32)
            q = (cand[ce] - cand[fl]) * (t - fl) + cand[fl]
            q = q * (ref != 0)
            q = q / _seq_sum(q, F64)
            keep = q != 0
            p = ref[keep]
            out[i] = _seq_sum(p * np.log(p / q[keep]), F64)
    return out


def kl_calibrate(data, levels, min_bins, bins, promotion="nep50"):
    """distribution_calibrate.py:117-171: first strict minimum wins, NaN never wins."""
    assert min_bins >= levels
    div = kl_divergences(data, levels, min_bins, bins, promotion)
    best, best_div = min_bins, np.inf
    for i in range(min_bins, bins):
        if div[i] < best_div:
            best_div, best = div[i], i
    return best


def kl_threshold(best_bins, fm_max, bins):
    """simulate_quantization.py:310 -> float32 input_max (regime independent)."""
    return F32((best_bins + 0.5) * (F64(F32(fm_max)) / bins))


# ---------------------------------------------------------------------------
# K6  export / QConv2D quantise
# ---------------------------------------------------------------------------
def quantize_int8_export(w, min_range, max_range):
    """MXNet contrib.quantize(out_type='int8') zero-centred (freeze.py:100-103)."""
    w = np.asarray(w, dtype=F32)
    real = F32(max(abs(F32(min_range)), abs(F32(max_range))))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        scale = F32(F32(127.0) / real)
        mag = np.minimum(((np.abs(w) * scale).astype(F32) + F32(0.5)).astype(F32), F32(127.0))
        out = (np.sign(w) * mag).astype(F32)
    return np.trunc(out).astype(np.int8), F32(-real), real


def qconv_quantize(x, min_range, max_range):
    """nn/quantized_conv.py:54-61 -> (int32 codes, scale)."""
    x = clip(x, min_range, max_range)
    min_range = F32(min_range)
    max_range = F32(max_range)
    if max_range == -min_range:
        scale = F32(max_range / F32(127))
    else:
        scale = F32(F32(max_range - min_range) / F32(255))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        codes = roundf((x / scale).astype(F32))
    return codes.astype(np.int32), scale


def qconv_quantize_auto(x, out_type="int8"):
    """nn/quantized_conv.py:63-72."""
    x = np.asarray(x, dtype=F32)
    if out_type == "int8":
        mx = F32(np.abs(x).max())
        mn = F32(-mx)
    elif out_type == "uint8":
        mn, mx = minmax(x)
    else:
        raise ValueError("unknown out type: ", out_type)
    return qconv_quantize(x, mn, mx)


def qconv_dequantize(y_int, scale):
    """nn/quantized_conv.py:74-76."""
    return (np.asarray(y_int).astype(F32) * F32(scale)).astype(F32)
