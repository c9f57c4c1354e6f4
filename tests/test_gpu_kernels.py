"""Parity tests proper: every kernel is called through the C ABI (ctypes -> libfq_b200.so) on the
GPU and compared BIT FOR BIT with the CPU oracle on the same seeded inputs, and with the golden
vectors the reference itself produced (tests/golden)."""
import os

import numpy as np
import pytest
import torch

from oracle import fq_oracle as O
from oracle import golden_recipes as R

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F32 = np.float32


@pytest.fixture(scope="module")
def ops():
    from quantization.mxnet_b200 import ops as _ops
    return _ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    if a.dtype.kind == "f":
        ai = a.view(np.uint32 if a.dtype == np.float32 else np.uint64)
        bi = b.view(np.uint32 if b.dtype == np.float32 else np.uint64)
        # +0.0 / -0.0 are both acceptable nowhere: require identical bits except NaN payloads
        same = (ai == bi) | (np.isnan(a) & np.isnan(b))
    else:
        same = a == b
    if not same.all():
        idx = np.argwhere(~same)[:5]
        raise AssertionError("mismatch at %s: got %s want %s (%d of %d differ)" % (
            idx.tolist(), a[tuple(idx.T)], b[tuple(idx.T)], (~same).sum(), same.size))


def rng(seed):
    return np.random.RandomState(seed)


# ---------------------------------------------------------------------------------------------
# K1
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,L", [(1, 1), (1, 1000), (1, 3_000_001), (128, 4096), (128, 50176), (1024, 9),
                                    (7, 4099), (64, 36864), (256, 64), (3, 2047), (3, 2048), (512, 4608)])
def test_absmax_rows(ops, rows, L):
    x = (rng(rows * 7 + L).standard_normal((rows, L)) * 3).astype(F32)
    got = host(ops.absmax_rows(dev(x), rows))
    bits_equal(got, O.absmax_rows(x, rows))
    # workspace invariant: a second call must give the same answer
    bits_equal(host(ops.absmax_rows(dev(x), rows)), O.absmax_rows(x, rows))


def test_absmax_unaligned_view(ops):
    x = rng(5).standard_normal(100_003).astype(F32)
    t = dev(x)[3:]          # 12-byte offset: scalar head path
    bits_equal(host(ops.absmax_rows(t, 1)), O.absmax_rows(x[3:], 1))


@pytest.mark.parametrize("n", [1, 5, 4096, 1_000_003])
def test_minmax(ops, n):
    x = rng(n).standard_normal(n).astype(F32)
    lo, hi = O.minmax(x)
    bits_equal(host(ops.minmax(dev(x))), np.array([lo, hi], dtype=F32))


@pytest.mark.parametrize("n", [1, 2, 128, 256, 1000, 5000])
def test_mean_kahan(ops, n):
    v = np.abs(rng(n).standard_normal(n) * 10).astype(F32)
    bits_equal(host(ops.mean_kahan(dev(v))), np.array([O.mean_kahan_f32(v)], dtype=F32))


@pytest.mark.parametrize("shape", [(128, 16, 32, 32), (128, 64, 8, 8), (256, 64), (128, 1024), (2, 3, 5, 5),
                                   (32, 32, 56, 56), (5, 7, 11, 13)])
def test_input_range(ops, shape):
    x = (rng(sum(shape)).standard_normal(shape) * 2).astype(F32)
    cur, per = O.input_range(x)
    ps = torch.empty(shape[0], dtype=torch.float32, device="cuda")
    got = ops.input_range(dev(x), per_sample=ps)
    bits_equal(host(got), np.array([cur], dtype=F32))
    bits_equal(host(ps), per)


@pytest.mark.parametrize("promotion", ["legacy", "nep50"])
@pytest.mark.parametrize("signed", [False, True])
def test_scale_from_max(ops, promotion, signed):
    r = rng(11)
    maxes = np.concatenate([np.abs(r.standard_normal(300)).astype(F32) * F32(7), np.array([0, 1e-38, 1e-30, 255, 6, 1e20], F32)])
    for bits in (2, 3, 4, 5, 6, 8, 12, 16):
        for lo_mode, layer in ((ops.LO_NEG_MAX if signed else ops.LO_ZERO, "conv"), (ops.LO_ZERO, "dense")):
            for m in maxes[:: 7 if bits not in (4, 8) else 1]:
                want = np.array(O.input_qparams(m, bits, signed, promotion, layer), dtype=F32)
                got = host(ops.scale_from_max(dev(np.array([m], F32)), bits, signed, lo_mode, promotion=promotion))
                bits_equal(got, want)


# ---------------------------------------------------------------------------------------------
# K2
# ---------------------------------------------------------------------------------------------
def tie_heavy(seed, n, d):
    """Values exactly on and next to rounding ties (k + 0.5) * d, plus random ones."""
    r = rng(seed)
    k = r.randint(-300, 300, n).astype(F32)
    x = ((k + F32(0.5)) * F32(d)).astype(F32)
    x[::3] = np.nextafter(x[::3], F32(np.inf))
    x[1::3] = np.nextafter(x[1::3], F32(-np.inf))
    x[::5] = (r.standard_normal(len(x[::5])) * 100 * d).astype(F32)
    return x


@pytest.mark.parametrize("bits,signed", [(8, False), (8, True), (4, False), (4, True), (2, True), (3, False),
                                         (16, False), (16, True), (6, True), (5, False), (12, True)])
@pytest.mark.parametrize("promotion", ["legacy", "nep50"])
def test_forward_scalar_device_qparams(ops, bits, signed, promotion):
    n = 70_001
    for max_ in (F32(2.5), F32(0.0), F32(1e-30), F32(6.0), F32(317.77)):
        d, s, lo, hi = O.input_qparams(max_, bits, signed, promotion)
        x = tie_heavy(bits, n, float(d) if d > 0 else 1.0)
        x[:4] = [max_, -max_, 0.0, -0.0]
        y, code = O.fake_quant_scalar(x, d, s, lo, hi)
        qp = ops.scale_from_max(dev(np.array([max_], F32)), bits, signed, ops.LO_NEG_MAX if signed else ops.LO_ZERO,
                                promotion=promotion)
        cdt = torch.int32
        gy, gc = ops.forward_scalar(dev(x), qp, codes_dtype=cdt)
        bits_equal(host(gy), y)
        assert np.array_equal(host(gc), code.astype(np.int32))


@pytest.mark.parametrize("cdt,bits,signed", [(torch.int8, 8, True), (torch.uint8, 8, False), (torch.int16, 16, True),
                                             (torch.uint16, 16, False), (torch.float32, 8, False)])
def test_forward_scalar_code_dtypes(ops, cdt, bits, signed):
    x = (rng(3).standard_normal(40_000) * 2).astype(F32)
    d, s, lo, hi = O.input_qparams(F32(2.0), bits, signed, "legacy")
    y, code = O.fake_quant_scalar(x, d, s, lo, hi)
    gy, gc = ops.forward_scalar_host(dev(x), float(d), float(s), float(lo), float(hi), True, codes_dtype=cdt)
    bits_equal(host(gy), y)
    gcn = gc.cpu().view(torch.int16).numpy().view(np.uint16) if cdt == torch.uint16 else host(gc)
    assert np.array_equal(gcn.astype(np.int64), code.astype(np.int64))


def test_forward_scalar_noclip_and_unaligned(ops):
    x = (rng(4).standard_normal(10_007) * 5).astype(F32)
    y, _ = O.fake_quant_scalar(x, F32(0.0173), F32(0.0172))
    bits_equal(host(ops.forward_scalar_host(dev(x), 0.0173, 0.0172, clip=False)), y)
    t = dev(x)[1:]           # misaligned input -> scalar kernel
    y2, _ = O.fake_quant_scalar(x[1:], F32(0.0173), F32(0.0172), F32(-1), F32(1))
    bits_equal(host(ops.forward_scalar_host(t, 0.0173, 0.0172, -1.0, 1.0, True)), y2)


def test_forward_scalar_empty(ops):
    x = torch.empty(0, dtype=torch.float32, device="cuda")
    assert ops.forward_scalar_host(x, 1.0, 1.0).numel() == 0


@pytest.mark.parametrize("rows,L", [(1, 5000), (64, 36864), (1024, 9), (512, 4608), (32, 1027), (10, 64), (7, 4099)])
@pytest.mark.parametrize("bits", [8, 4, 2])
def test_forward_rows(ops, rows, L, bits):
    w = (rng(rows + L + bits).standard_normal((rows, L)) * 0.1).astype(F32)
    w[0, 0] = 0
    s, d, _ = O.weight_scales(w, rows, bits)
    y, code = O.fake_quant_rows(w, rows, s, d)
    gy, gc = ops.forward_rows(dev(w), dev(s), codes_dtype=torch.int8)
    bits_equal(host(gy), y)
    assert np.array_equal(host(gc), code.astype(np.int8))


@pytest.mark.parametrize("shape", [(128, 16, 32, 32), (128, 64), (4, 3, 5, 5), (64, 32, 28, 28), (16, 96, 56, 56)])
@pytest.mark.parametrize("bits,signed,layer", [(8, False, "conv"), (8, True, "conv"), (4, True, "dense"), (4, False, "conv")])
@pytest.mark.parametrize("promotion", ["legacy", "nep50"])
def test_forward_online(ops, shape, bits, signed, layer, promotion):
    x = (rng(sum(shape) + bits).standard_normal(shape) * 1.7).astype(F32)
    if not signed:
        x = np.maximum(x, 0)
    y, code, cur, qp = O.fake_quant_input(x, bits, signed, None, promotion, layer)
    lo_mode = ops.LO_NEG_MAX if (signed and layer == "conv") else ops.LO_ZERO
    gy, gcur, gqp, gc = ops.forward_online(dev(x), bits, signed, lo_mode, promotion=promotion, codes_dtype=torch.int32)
    bits_equal(host(gcur), np.array([cur], F32))
    bits_equal(host(gqp), np.array(qp, F32))
    bits_equal(host(gy), y)
    assert np.array_equal(host(gc), code.astype(np.int32))
    # repeat: the workspace must have been left clean
    gy2, _, _ = ops.forward_online(dev(x), bits, signed, lo_mode, promotion=promotion)
    bits_equal(host(gy2), y)


@pytest.mark.parametrize("shape", [(128, 16, 32, 32), (128, 64), (8, 32, 56, 56)])
def test_forward_online_offline_range_and_tracking(ops, shape):
    x = np.maximum(rng(9).standard_normal(shape), 0).astype(F32)
    imax = F32(1.25)
    y, code, cur, qp = O.fake_quant_input(x, 8, False, imax, "legacy", "conv")
    gy, gcur, gqp = ops.forward_online(dev(x), 8, False, ops.LO_ZERO, input_max=dev(np.array([imax], F32)))
    bits_equal(host(gy), y)
    bits_equal(host(gcur), np.array([cur], F32))
    bits_equal(host(gqp), np.array(qp, F32))
    # range tracking only (quantize_input switched off, convert_conv2d.py:55-57)
    gy, gcur, _ = ops.forward_online(dev(x), 8, False, ops.LO_ZERO, quantize=False)
    assert gy is None
    bits_equal(host(gcur), np.array([cur], F32))


WEIGHT_SHAPES = [(64, 3, 7, 7), (32, 1, 3, 3), (512, 512, 3, 3), (1000, 1024), (16, 16, 1, 1), (10, 64), (256, 64, 1, 1)]


@pytest.mark.parametrize("shape", WEIGHT_SHAPES)
@pytest.mark.parametrize("quant_type", ["layer", "channel"])
@pytest.mark.parametrize("bits", [8, 4])
def test_quant_weight(ops, shape, quant_type, bits):
    w = (rng(sum(shape)).standard_normal(shape) * 0.05).astype(F32)
    y, code, s = O.fake_quant_weight(w, bits, quant_type)
    rows = shape[0] if quant_type == "channel" else 1
    gy, _, gs, gc = ops.quant_weight(dev(w), rows, bits, codes_dtype=torch.int8)
    bits_equal(host(gs), s)
    bits_equal(host(gy), y)
    assert np.array_equal(host(gc), code.astype(np.int8))
    gy2, _, _ = ops.quant_weight(dev(w), rows, bits)
    bits_equal(host(gy2), y)


@pytest.mark.parametrize("shape", [(64, 3, 7, 7), (32, 1, 3, 3), (256, 256, 3, 3), (128, 64, 1, 1)])
@pytest.mark.parametrize("rows_kind", ["layer", "channel", "fold_only"])
@pytest.mark.parametrize("with_bias", [False, True])
def test_quant_weight_with_bn_fold(ops, shape, rows_kind, with_bias):
    r = rng(sum(shape) + 1)
    cout = shape[0]
    w = (r.standard_normal(shape) * 0.05).astype(F32)
    gamma = (1 + 0.3 * r.standard_normal(cout)).astype(F32)
    beta = (0.1 * r.standard_normal(cout)).astype(F32)
    mean = (0.2 * r.standard_normal(cout)).astype(F32)
    var = np.abs(1 + 0.5 * r.standard_normal(cout)).astype(F32)
    var[0] = 0.0                                   # eps 1e-10 path
    bias = (0.05 * r.standard_normal(cout)).astype(F32) if with_bias else None
    w2, b2 = O.fold_bn(w, bias, gamma, beta, mean, var)
    args = dict(gamma=dev(gamma), beta=dev(beta), mean=dev(mean), var=dev(var), bias=None if bias is None else dev(bias))
    if rows_kind == "fold_only":
        gy, gb, _ = ops.quant_weight(dev(w), 1, 0, **args)
        bits_equal(host(gy), w2)
        bits_equal(host(gb), b2)
        return
    qt = rows_kind
    y, code, s = O.fake_quant_weight(w2, 4, qt)
    rows = cout if qt == "channel" else 1
    gy, gb, gs = ops.quant_weight(dev(w), rows, 4, **args)
    bits_equal(host(gb), b2)
    bits_equal(host(gs), s)
    bits_equal(host(gy), y)


def test_quant_weight_group_rows(ops):
    # rows = G with 1 < G < Cout: an extension (the reference's broadcast only allows G in {1, Cout})
    w = (rng(8).standard_normal((32, 4, 3, 3)) * 0.1).astype(F32)
    s, d, _ = O.weight_scales(w, 4, 8)
    y, _ = O.fake_quant_rows(w, 4, s, d)
    gy, _, gs = ops.quant_weight(dev(w), 4, 8)
    bits_equal(host(gs), s)
    bits_equal(host(gy), y)


# ---------------------------------------------------------------------------------------------
# K3 / K4
# ---------------------------------------------------------------------------------------------
def test_ste_backward(ops):
    x = (rng(1).standard_normal(50_001) * 2).astype(F32)
    dy = rng(2).standard_normal(50_001).astype(F32)
    t = dev(dy)
    assert ops.ste_backward(t) is t                         # identity: aliased, as ste_func.py:43-44
    qp = dev(np.array([0.01, 0.01, -1.5, 1.5], F32))
    got = host(ops.ste_backward(t, dev(x), qp, mode=ops.STE_CLIP_MASK))
    bits_equal(got, O.ste_backward(dy, x, F32(-1.5), F32(1.5), "mask"))


@pytest.mark.parametrize("promotion", ["legacy", "nep50"])
def test_ema_update(ops, promotion):
    r = rng(6)
    state = np.abs(r.standard_normal(53)).astype(F32)
    state[:3] = 0
    for step in range(5):
        cur = np.abs(r.standard_normal(53) * 3).astype(F32)
        want = O.ema_scalar(state, cur, 0.9, promotion)
        t = dev(state)
        ops.ema_update(t, dev(cur), 0.9, True, promotion=promotion)
        bits_equal(host(t), want)
        state = want
    cur = r.standard_normal(2048).astype(F32)
    st = r.standard_normal(2048).astype(F32)
    t = dev(st)
    ops.ema_update(t, dev(cur), 0.9, False)
    bits_equal(host(t), O.ema_tensor(st, cur, 0.9))


# ---------------------------------------------------------------------------------------------
# K5
# ---------------------------------------------------------------------------------------------
def gpu_hist_batches(ops, batches, bins, promotion):
    counts = torch.zeros(bins + 1, dtype=torch.int64, device="cuda")
    hist = torch.zeros(bins + 1, dtype=torch.float32, device="cuda")
    seen = torch.zeros(1, dtype=torch.int32, device="cuda")
    mx = None
    per_batch = []
    for b, fm in enumerate(batches):
        t = dev(fm)
        if mx is None:
            mx = ops.minmax(t)[1:2].clone()
        ops.hist_nonzero(t, mx, bins, counts, promotion=promotion)
        per_batch.append(host(counts).copy())
        ops.hist_accumulate(counts, hist, b == 0, seen)
        assert int(counts.abs().sum()) == 0
    return per_batch, host(hist), host(mx)[0], int(seen[0])


@pytest.mark.parametrize("name", sorted(R.hist_cases().keys()))
def test_histogram_vs_reference_golden(ops, name):
    g = np.load(os.path.join(GOLD, "hist_nep50.npz"))
    batches = R.hist_cases()[name]
    if str(g["hist/%s/error" % name]):
        batches = batches[:1]
    per_batch, hist, mx, seen = gpu_hist_batches(ops, batches, R.BINS, "nep50")
    assert F32(mx) == g["hist/%s/max" % name]
    for b in range(len(batches)):
        want = g["hist/%s/batch%d" % (name, b)]
        got = per_batch[b]
        assert np.array_equal(got[: len(want)], want.astype(np.int64)) and got[len(want):].sum() == 0
    if not str(g["hist/%s/error" % name]):
        want = g["hist/%s/acc" % name]
        bits_equal(hist[: len(want)], want)
        assert seen == (len(want) == R.BINS + 1)


@pytest.mark.parametrize("promotion", ["legacy", "nep50"])
def test_histogram_vs_oracle_both_regimes(ops, promotion):
    for seed, sc in ((1, 1.0), (2, 37.3), (3, 300.0), (4, 1e-3)):
        x = R.relu_normal(seed, 200_003, sc)
        want = O.histogram_counts(x, R.BINS, x.max(), promotion)
        per_batch, _, _, _ = gpu_hist_batches(ops, [x], R.BINS, promotion)
        assert np.array_equal(per_batch[0][: len(want)], want)


def test_histogram_multi_tensor_equals_single(ops):
    shapes = [(128, 32, 14, 14), (128, 1024), (7, 3, 5, 5), (64, 64, 28, 28), (1,), (128, 256, 7, 7)] * 12   # 72 > FQ_MAX_BATCH
    xs = [dev(R.relu_normal(40 + i, int(np.prod(s)), 1.0 + 0.5 * i).reshape(s)) for i, s in enumerate(shapes)]
    minmax = torch.stack([ops.minmax(x) for x in xs])
    minmax[3, 1] *= 0.5                                   # a frozen max smaller than the data: clipping
    want = torch.zeros(len(xs), R.BINS + 1, dtype=torch.int64, device="cuda")
    for i, x in enumerate(xs):
        ops.hist_nonzero(x, minmax[i, 1:2], R.BINS, want[i])
    got = torch.zeros_like(want)
    ops.hist_nonzero_multi(xs, minmax, 2, 1, R.BINS, got)
    assert torch.equal(got, want)
    assert int(got.sum()) == sum(int((x != 0).sum()) for x in xs)


@pytest.mark.parametrize("bins", [1, 7, 100, 2048, 4096, 8192])
def test_histogram_other_bin_counts_ragged_and_unaligned(ops, bins):
    r = rng(bins)
    for n, off in ((1, 0), (3, 0), (4099, 0), (70_001, 1), (300_000, 3), (1_234_567, 0)):
        x = np.maximum(r.standard_normal(n + off) * 3, 0).astype(F32)
        x[::7] = 0
        t = dev(x)[off:]                       # off != 0: not 16 B aligned -> scalar kernel
        for promo in ("legacy", "nep50"):
            mx = F32(x[off:].max() * 0.8) if n > 3 else F32(max(x[off:].max(), 0.5))     # clipping at a frozen max
            want = O.histogram_counts(x[off:], bins, mx, promo)
            counts = torch.zeros(bins + 1, dtype=torch.int64, device="cuda")
            ops.hist_nonzero(t, dev(np.array([mx], F32)), bins, counts, promotion=promo)
            got = host(counts)
            assert np.array_equal(got[:len(want)], want) and got[len(want):].sum() == 0, (n, off, promo)


def test_histogram_without_a_positive_max_counts_nothing(ops):
    """The reference asserts max_ > 0 before it histograms (distribution_calibrate.py:36); the kernel leaves the
    counters alone and the deferred assert of collect_feature_maps reports it."""
    x = dev(np.abs(rng(1).standard_normal(10_000)).astype(F32))
    for mx in (0.0, -1.0, float("nan")):
        counts = torch.zeros(2049, dtype=torch.int64, device="cuda")
        ops.hist_nonzero(x, torch.tensor([mx], device="cuda"), 2048, counts)
        assert int(counts.sum()) == 0
    # NaN / negative / -0.0 elements are dropped, +inf lands in the last bin
    y = dev(np.array([np.nan, -1.0, -0.0, 0.0, np.inf, 0.5, 1.0], F32))
    counts = torch.zeros(2049, dtype=torch.int64, device="cuda")
    ops.hist_nonzero(y, torch.tensor([1.0], device="cuda"), 2048, counts)
    want = O.histogram_counts(np.array([0.0, 0.0, 1.0, 0.5, 1.0], F32), 2048, F32(1.0), "legacy")
    got = host(counts)
    assert np.array_equal(got[:len(want)], want) and got[len(want):].sum() == 0


def test_hist_accumulate_many_batches_in_one_launch_replays_the_adds_in_order(ops):
    """counts [steps, n]: one float32 add per batch in batch order (distribution_calibrate.py:47,103-104),
    the same bits as one launch per batch (what a deferred all-reduce of many batches relies on)."""
    r = np.random.RandomState(3)
    steps, n = 9, 3 * 2049
    c = r.randint(0, 1 << 27, size=(steps, n)).astype(np.int64)       # float32(count) rounds
    c[:, -1] = 0
    want = np.zeros(n, F32)
    for s in range(steps):
        want = c[s].astype(F32) if s == 0 else want + c[s].astype(F32)
    for first in (True, False):
        counts = dev(c)
        hist = torch.full((n,), 5.0, device="cuda")
        seen = torch.zeros(1, dtype=torch.int32, device="cuda")
        ops.hist_accumulate(counts.view(-1), hist, first, seen)
        w = want
        if not first:
            w = np.full(n, 5.0, F32)
            for s in range(steps):
                w = w + c[s].astype(F32)
        bits_equal(host(hist), w)
        assert int(counts.abs().sum()) == 0 and int(seen[0]) == 0
    counts = dev(c)
    counts[4, -1] = 7
    seen = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.hist_accumulate(counts.view(-1), torch.zeros(n, device="cuda"), True, seen)
    assert int(seen[0]) == 1
    with pytest.raises(Exception):
        ops.hist_accumulate(torch.zeros(n + 1, dtype=torch.int64, device="cuda"), torch.zeros(n, device="cuda"), True)


_KL = [(n, l) for n, ls in R.KL_LEVELS.items() for l in ls]


@pytest.mark.parametrize("name,levels", _KL)
def test_kl_search_vs_reference_golden(ops, name, levels):
    g = np.load(os.path.join(GOLD, "kl_nep50.npz"))
    h = g["kl/%s/hist" % name]
    best, div = ops.kl_search(dev(h), levels, levels, R.BINS, promotion="nep50")
    assert int(best[0]) == int(g["kl/%s/L%d/best" % (name, levels)])
    want = O.kl_divergences(h, levels, levels, R.BINS, "nep50")
    got = host(div)[0]
    ok = np.isclose(got[levels:], want[levels:], rtol=1e-11, atol=1e-13, equal_nan=True)
    assert ok.all(), (np.argwhere(~ok)[:5].tolist(), got[levels:][~ok][:5], want[levels:][~ok][:5])


@pytest.mark.parametrize("name", ["relu", "huge_counts", "lognormal", "non_integer", "len2049"])
def test_kl_search_legacy_regime_vs_oracle(ops, name):
    h = R.kl_hist_cases()[name]
    levels = R.KL_LEVELS[name][0]
    best, div = ops.kl_search(dev(h), levels, levels, R.BINS, promotion="legacy")
    want = O.kl_divergences(h, levels, levels, R.BINS, "legacy")
    assert int(best[0]) == O.kl_calibrate(h, levels, levels, R.BINS, "legacy")
    assert np.isclose(host(div)[0][levels:], want[levels:], rtol=1e-11, atol=1e-13, equal_nan=True).all()


def test_kl_search_batched_windows_and_threshold(ops):
    g = np.load(os.path.join(GOLD, "kl_nep50.npz"))
    names = ["relu", "outliers", "lognormal"]
    hs = np.stack([g["kl/%s/hist" % n] for n in names])
    best, _ = ops.kl_search(dev(hs), 256, 256, R.BINS, promotion="nep50")
    want = [int(g["kl/%s/L256/best" % n]) for n in names]
    assert host(best).tolist() == want
    mx = np.array([4.2, 37.5, 0.013], F32)
    th = host(ops.kl_threshold(best, dev(mx), R.BINS))
    bits_equal(th, np.array([O.kl_threshold(b, m, R.BINS) for b, m in zip(want, mx)], F32))


def test_kl_search_rejects_bad_arguments(ops):
    from quantization.mxnet_b200._ffi import FQError
    h = torch.ones(2048, device="cuda")
    with pytest.raises(FQError, match="min_bins should be greater than levels"):
        ops.kl_search(h, 256, 128, 2048)
    with pytest.raises(FQError, match="no CPU path"):
        ops.absmax_rows(torch.ones(8), 1)


# ---------------------------------------------------------------------------------------------
# K6
# ---------------------------------------------------------------------------------------------
def test_int8_export_and_qconv(ops):
    w = (rng(12).standard_normal((64, 32, 3, 3)) * 0.2).astype(F32)
    mx = F32(np.abs(w).max())
    q, lo, hi = O.quantize_int8_export(w, -mx, mx)
    gq, gr = ops.quantize_int8_export(dev(w), dev(np.array([-mx, mx], F32)))
    assert np.array_equal(host(gq), q)
    bits_equal(host(gr), np.array([lo, hi], F32))

    x = rng(13).uniform(size=(2, 2, 7, 7)).astype(F32)
    for out_type in ("int8", "uint8"):
        codes, scale = O.qconv_quantize_auto(x, out_type)
        rg = np.array([-np.abs(x).max(), np.abs(x).max()], F32) if out_type == "int8" else np.array(O.minmax(x), F32)
        gc, gs = ops.qconv_quantize(dev(x), dev(rg))
        assert np.array_equal(host(gc), codes)
        bits_equal(host(gs), np.array([scale], F32))
    acc = rng(14).randint(-2 ** 20, 2 ** 20, (4, 8, 5, 5)).astype(np.int32)
    s_in, s_w = F32(0.0123), F32(0.00071)
    got = host(ops.qconv_dequantize(dev(acc), dev(np.array([s_in], F32)), dev(np.array([s_w], F32))))
    bits_equal(got, O.qconv_dequantize(acc, F32(s_in * s_w)))


# ---------------------------------------------------------------------------------------------
# full-size properties (BASELINE config sizes; the oracle would take too long there)
# ---------------------------------------------------------------------------------------------
def test_full_size_properties(ops):
    n_samples, chw = 128, 64 * 112 * 112            # largest MobileNet-1.0 layer input: 102,760,448 elements
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(n_samples, chw, device="cuda", generator=g).clamp_(min=0)
    y, cur, qp, codes = ops.forward_online(x, 8, False, ops.LO_ZERO, codes_dtype=torch.uint8)
    per = x.abs().amax(dim=1)
    assert torch.equal(ops.absmax_rows(x, n_samples), per)
    assert abs(float(cur) - float(per.double().mean())) < 1e-5 * float(cur)
    # idempotence: quantising a quantised tensor with the same qparams changes nothing
    y2 = ops.forward_scalar(y, qp)
    assert torch.equal(y, y2)
    # 8-bit codes that reproduce the output: y == code * s, all inside [0, hi]
    assert y.min() >= 0 and float(y.max()) <= float(qp[3]) * (1 + 1e-6)
    assert torch.equal(y, codes.float() * qp[1])
    # histogram: checksum of counts == number of clipped non-zero elements
    counts = torch.zeros(R.BINS + 1, dtype=torch.int64, device="cuda")
    mx = ops.minmax(x)[1:2].clone()
    ops.hist_nonzero(x, mx, R.BINS, counts)
    assert int(counts.sum()) == int((x != 0).sum())
    # linearity: hist(x) + hist(x) == 2 * hist(x)
    c1 = counts.clone()
    ops.hist_nonzero(x, mx, R.BINS, counts)
    assert torch.equal(counts, 2 * c1)


def test_beyond_2_31_elements_uses_64_bit_indexing(ops):
    """BASELINE config 5 reaches 2^32 fp32 elements; anything indexed with 32 bits breaks past 2^31."""
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs ~20 GB of device memory")
    n = (1 << 31) + 3 * 4096 + 8
    x = torch.empty(n, device="cuda")
    x.uniform_(0, 1)
    x[-1] = 7.5                        # the maximum sits at the very last element
    x[(1 << 31) + 5] = 0.0
    assert float(ops.absmax_rows(x, 1)) == 7.5
    mm = ops.minmax(x)
    assert float(mm[1]) == 7.5 and float(mm[0]) == 0.0
    qp = ops.scale_from_max(mm[1:2].clone(), 8, False, ops.LO_ZERO)
    y = ops.forward_scalar(x, qp)
    assert F32(y[-1].item()) == F32(255) * F32(qp[1].item()) and float(y[(1 << 31) + 5]) == 0.0
    tail = slice(n - 100_000, n)
    want, _ = O.fake_quant_scalar(host(x[tail]), *[F32(v) for v in host(qp)])
    bits_equal(host(y[tail]), want)
    counts = torch.zeros(R.BINS + 1, dtype=torch.int64, device="cuda")
    ops.hist_nonzero(x, mm[1:2].clone(), R.BINS, counts)
    del y
    nz = sum(int((x[i:i + (1 << 28)] != 0).sum()) for i in range(0, n, 1 << 28))
    assert int(counts.sum()) == nz
    assert int(counts[R.BINS - 1] + counts[R.BINS]) >= 1          # the 7.5 landed in the top bin


def mean_close(got, want, y):
    """1 ULP of the result + 2^-23 of mean |y| (Kahan's bound is relative to sum |y|: see test_channel_stats_math)."""
    mag = np.abs(y.astype(np.float64)).mean(axis=(0, 2, 3))
    tol = np.spacing(np.abs(want)).astype(np.float64) + 2.0 ** -23 * mag
    return bool(np.all(np.abs(got.astype(np.float64) - want.astype(np.float64)) <= tol))


def ulp_diff(a, b):
    """Largest distance in units of the last place between two float32 arrays of one sign pattern."""
    a = np.ascontiguousarray(a, F32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, F32).view(np.int32).astype(np.int64)
    return int(np.abs(a - b).max())


@pytest.mark.parametrize("shape,offset", [((8, 16, 14, 14), 0.5), ((128, 64, 8, 8), 0.5), ((4, 3, 7, 7), 0.5),
                                          ((32, 256, 1, 1), 0.5), ((2, 2048, 7, 7), 0.5), ((16, 5, 13, 11), 0.5),
                                          ((1, 8, 32, 32), 0.5), ((16, 8, 56, 56), 100.0), ((64, 32, 28, 28), -3.0),
                                          ((3, 4, 33, 33), 1e4)])
def test_channel_stats_fake_bn(ops, shape, offset):
    """convert_conv2d.py:150-153.  One-pass float64 shifted moments vs the oracle's two-pass sequential Kahan fp32
    sums: both sit within ~1 ULP of the exact value, so they agree to <= 2 ULP (measured: 0-1)."""
    from oracle import build_c as C
    y = (rng(sum(shape)).standard_normal(shape) * 2 + offset).astype(F32)
    parts = torch.empty(shape[1], 4, dtype=torch.float64, device="cuda")
    mean, var = ops.channel_stats(dev(y), parts=parts)
    want_m, want_v = C.channel_stats(y)
    assert mean_close(host(mean), want_m, y), ulp_diff(host(mean), want_m)
    assert ulp_diff(host(var), want_v) <= 2, ulp_diff(host(var), want_v)
    m2, v2 = ops.channel_stats(dev(y))            # deterministic, and the workspace is left clean
    bits_equal(host(m2), host(mean))
    bits_equal(host(v2), host(var))
    bits_equal(host(ops.absmax_rows(dev(y), shape[0])), O.absmax_rows(y, shape[0]))
    # the records: n, and K = the channel's first element
    rec = host(parts)
    assert np.all(rec[:, 0] == shape[0] * shape[2] * shape[3])
    assert np.array_equal(rec[:, 3].astype(F32), y[0, :, 0, 0])
    # a constant channel has exactly zero variance and its own value as mean, whatever its magnitude
    yc = np.full((4, 3, 9, 9), 1234.5678, F32)
    mc, vc = ops.channel_stats(dev(yc))
    wm, wv = C.channel_stats(yc)
    bits_equal(host(mc), wm)
    assert float(host(vc).max()) <= float(wv.max()) + 1e-12


@pytest.mark.parametrize("ranks", [2, 4, 8])
def test_channel_stats_records_of_shards_reproduce_the_global_batch(ops, ranks):
    """Data parallel: every rank's {n, S1, S2, K} records, all-gathered and combined by fq_channel_stats_finish,
    give the statistics of the global batch to the single-GPU bound."""
    from oracle import build_c as C
    shape = (16 * ranks, 24, 14, 14)
    y = (rng(ranks).standard_normal(shape) * 1.5 + rng(ranks + 1).standard_normal((1, 24, 1, 1)) * 4).astype(F32)
    recs = torch.empty(ranks, 24, 4, dtype=torch.float64, device="cuda")
    for r in range(ranks):
        ops.channel_stats(dev(y[16 * r:16 * (r + 1)]), parts=recs[r], finish=False)
    mean, var = ops.channel_stats_finish(recs)
    want_m, want_v = C.channel_stats(y)
    one_m, one_v = ops.channel_stats(dev(y))
    assert mean_close(host(mean), want_m, y) and ulp_diff(host(var), want_v) <= 2
    assert mean_close(host(mean), host(one_m), y) and ulp_diff(host(var), host(one_v)) <= 1


def test_randomised_shapes_against_oracle(ops):
    """Seeded fuzzing over ragged shapes, bit widths, signedness and promotion regimes: every slice / tile /
    short-row / long-row code path is hit with sizes that are not multiples of anything."""
    r = rng(2024)
    for case in range(60):
        n_s = int(r.choice([1, 2, 3, 7, 16, 33, 128]))
        L = int(r.choice([1, 5, 9, 31, 64, 257, 1024, 2047, 2048, 2051, 4096, 4100, 9001, 40000]))
        bits = int(r.choice([2, 3, 4, 5, 6, 8, 12, 16]))
        signed = bool(r.randint(2))
        promo = ["legacy", "nep50"][r.randint(2)]
        x = (r.standard_normal((n_s, L)) * float(r.choice([1e-3, 1.0, 40.0]))).astype(F32)
        if not signed:
            x = np.maximum(x, 0)
        layer = ["conv", "dense"][r.randint(2)]
        lo_mode = ops.LO_NEG_MAX if (signed and layer == "conv") else ops.LO_ZERO
        # online
        y, code, cur, qp = O.fake_quant_input(x, bits, signed, None, promo, layer)
        gy, gcur, gqp = ops.forward_online(dev(x), bits, signed, lo_mode, promotion=promo)
        bits_equal(host(gcur), np.array([cur], F32))
        bits_equal(host(gy), y)
        # offline range + tracking
        imax = F32(abs(r.standard_normal()) + 0.1)
        y, code, cur, qp = O.fake_quant_input(x, bits, signed, imax, promo, layer)
        gy, gcur, gqp = ops.forward_online(dev(x), bits, signed, lo_mode, input_max=dev(np.array([imax], F32)), promotion=promo)
        bits_equal(host(gcur), np.array([cur], F32))
        bits_equal(host(gqp), np.array(qp, F32))
        bits_equal(host(gy), y)
        # weights, per-row and per-layer, treating x as [rows, L]
        w = (x - x.mean()).astype(F32)
        for rows in (n_s, 1):
            s, d, _ = O.weight_scales(w, rows, max(bits, 2))
            wy, _ = O.fake_quant_rows(w, rows, s, d)
            gw, _, gs = ops.quant_weight(dev(w), rows, max(bits, 2))
            bits_equal(host(gs), s)
            bits_equal(host(gw), wy)
            bits_equal(host(ops.forward_rows(dev(w), dev(s))), wy)
        # histogram
        xa = np.abs(x)
        if xa.max() > 0:
            want = O.histogram_counts(xa, R.BINS, xa.max(), promo)
            counts = torch.zeros(R.BINS + 1, dtype=torch.int64, device="cuda")
            ops.hist_nonzero(dev(xa), dev(np.array([xa.max()], F32)), R.BINS, counts, promotion=promo)
            assert np.array_equal(host(counts)[:len(want)], want)


def test_c_abi_rejects_bad_arguments_with_messages(ops):
    """Error behaviour of the boundary: non-zero return + fq_last_error(), surfaced as FQError; nothing silently
    falls back or launches on bad input."""
    from quantization.mxnet_b200._ffi import FQError
    x = torch.randn(8, 16, device="cuda")
    with pytest.raises(FQError, match="contiguous"):
        ops._ffi.dl(x.t())
    with pytest.raises(FQError, match="float32"):
        ops.forward_scalar_host(x.double(), 1.0, 1.0)
    with pytest.raises(FQError, match="not divisible|divisible"):
        ops.absmax_rows(x, 3)
    with pytest.raises(FQError, match="rows"):
        ops.absmax_rows(x, 0)
    with pytest.raises(FQError, match="bits"):
        ops.scale_from_max(torch.ones(1, device="cuda"), 1, False, ops.LO_ZERO)
    with pytest.raises(FQError, match="qparams"):
        ops.forward_scalar(x, torch.ones(3, device="cuda"))
    with pytest.raises(FQError, match="counts"):
        ops.hist_nonzero(x.abs(), torch.ones(1, device="cuda"), 2048, torch.zeros(2048, dtype=torch.int64, device="cuda"))
    with pytest.raises(FQError, match="unsupported|codes dtype"):
        ops.forward_scalar_host(x, 1.0, 1.0, codes_dtype=torch.float64)
    with pytest.raises(FQError, match="given together"):
        ops.quant_weight(torch.randn(4, 4, device="cuda"), 1, 8, gamma=torch.ones(4, device="cuda"))
    with pytest.raises(FQError, match="divide Cout|must divide"):
        ops.quant_weight(torch.randn(6, 4, device="cuda"), 4, 8)
    # after all those failures the library still works and the workspace is clean
    bits_equal(host(ops.absmax_rows(x, 8)), O.absmax_rows(host(x), 8))


def test_quant_weight_multi_equals_per_tensor(ops):
    """One launch per phase for a whole network's weights == fq_quant_weight block by block, bit for bit."""
    r = rng(77)
    shapes = [(32, 3, 3, 3), (32, 1, 3, 3), (64, 32, 1, 1), (10, 64), (256, 256, 3, 3), (16, 16, 1, 1), (96, 1, 3, 3)] * 10
    jobs = []
    for i, shp in enumerate(shapes):          # 70 jobs: more than one launch's worth, mixed kinds
        w = dev((r.standard_normal(shp) * 0.1).astype(F32))
        kind = i % 4
        jb = {"w": w, "rows": shp[0] if kind in (0, 2) else 1, "bits": [8, 4, 2, 0][kind]}
        if kind in (2, 3) and len(shp) == 4:
            c = shp[0]
            jb.update(gamma=dev((1 + 0.2 * r.standard_normal(c)).astype(F32)), beta=dev((0.1 * r.standard_normal(c)).astype(F32)),
                      mean=dev((0.1 * r.standard_normal(c)).astype(F32)), var=dev(np.abs(1 + 0.3 * r.standard_normal(c)).astype(F32)),
                      bias=dev((0.05 * r.standard_normal(c)).astype(F32)) if i % 8 < 4 else None)
        elif kind == 3:
            jb["bits"] = 8
        jobs.append(jb)
    plan = ops.WeightPlan(jobs)
    for _ in range(2):                        # twice: the workspace must come back clean
        ws, bs, ss = ops.quant_weight_multi(plan)
        for i, jb in enumerate(jobs):
            fold = jb.get("gamma") is not None
            w1, b1, s1 = ops.quant_weight(jb["w"], jb["rows"], jb["bits"], jb.get("gamma"), jb.get("beta"), jb.get("mean"),
                                          jb.get("var"), jb.get("bias"))
            assert torch.equal(ws[i], w1), i
            if fold:
                assert torch.equal(bs[i], b1), i
            if jb["bits"] > 0:
                assert torch.equal(ss[i], s1), i
    bits_equal(host(ops.absmax_rows(jobs[0]["w"], 1)), O.absmax_rows(host(jobs[0]["w"]), 1))


def test_kl_search_many_levels_takes_the_block_per_candidate_kernel(ops):
    """levels = 1024 does not fit the 32-candidate kernel's shared memory: the by-the-book kernel runs."""
    h = R.kl_hist_cases()["lognormal"]
    best, div = ops.kl_search(dev(h), 1024, 1024, R.BINS, promotion="nep50")
    want = O.kl_divergences(h, 1024, 1024, R.BINS, "nep50")
    assert int(best[0]) == O.kl_calibrate(h, 1024, 1024, R.BINS, "nep50")
    assert np.isclose(host(div)[0][1024:], want[1024:], rtol=1e-11, atol=1e-13, equal_nan=True).all()


# ---------------------------------------------------------------------------------------------
# Winograd-domain weight quantisation (convert_conv2d.py:71-83)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["F23", "F43", "F63"])
@pytest.mark.parametrize("shape,bits", [((16, 16, 3, 3), 8), ((7, 5, 3, 3), 8), ((64, 1, 3, 3), 4), ((40, 33, 3, 3), 2),
                                        ((256, 128, 3, 3), 8)])
def test_wino_weight_quant_matches_oracle(ops, name, shape, bits):
    w = (rng(shape[0] * 13 + shape[1]).standard_normal(shape) * 0.1).astype(F32)
    w[0, 0] = 0                                # an all-zero kernel
    G, GI, GTI = O.winograd_matrices(name)
    want, want_s, _, _ = O.fake_quant_weight_wino(w, name, bits)
    got, got_s = ops.quant_weight_wino(dev(w), dev(G), dev(GI), dev(GTI), bits)
    bits_equal(host(got_s), want_s)
    bits_equal(host(got), want)
    # the workspace is restored: a second call gives the same answer
    got2, _ = ops.quant_weight_wino(dev(w), dev(G), dev(GI), dev(GTI), bits)
    bits_equal(host(got2), want)
    # and the plain weight path still works on the same workspace
    wq, _, sc = ops.quant_weight(dev(w), shape[0], bits)
    y, _, s = O.fake_quant_weight(w, bits, "channel")
    bits_equal(host(wq), y)


def test_wino_all_zero_weight_and_errors(ops):
    G, GI, GTI = (dev(m) for m in O.winograd_matrices("F43"))
    z = torch.zeros(4, 3, 3, 3, device="cuda")
    out, s = ops.quant_weight_wino(z, G, GI, GTI, 8)
    assert float(out.abs().max()) == 0.0 and float(s.abs().max()) == 0.0
    with pytest.raises(Exception, match="Cout, Cin, 3, 3"):
        ops.quant_weight_wino(torch.zeros(4, 3, 5, 5, device="cuda"), G, GI, GTI, 8)
    with pytest.raises(Exception, match="GI must be"):
        ops.quant_weight_wino(z, G, GTI, GI, 8)


@pytest.mark.parametrize("name", ["F23", "F43", "F63"])
def test_wino_backward_matches_oracle(ops, name):
    g = rng(17).standard_normal((33, 9, 3, 3)).astype(F32)
    G, GI, GTI = O.winograd_matrices(name)
    got = ops.wino_backward(dev(g), dev(G), dev(GI), dev(GTI))
    bits_equal(host(got), O.wino_backward(g, name))


def _near_tie_histograms():
    """Two histograms that differ in ONE ULP of ONE bin and whose KL arg-mins are 259 bins apart: a mix of the
    'relu' and 'lognormal' recipes at the point where the best candidate switches from 1792 to 1533, fine-tuned on
    bin 1063 by bisection against the oracle (found offline; the oracle re-checks both below)."""
    cases = R.kl_hist_cases()
    hA, hB = cases["relu"].astype(np.float64), cases["lognormal"].astype(np.float64)
    hB = hB * (hA.sum() / hB.sum())
    t = 0.5085390468650866
    h = ((1 - t) * hA + t * hB).astype(F32)
    out = []
    for bits in (0x40dc07d8, 0x40dc07d7):
        g = h.copy()
        g[1063] = np.uint32(bits).view(F32)
        out.append(g)
    return out


def test_kl_search_reports_the_margin_of_a_constructed_near_tie(ops):
    """north_star: chosen KL bins bit-exact.  The divergences agree with NumPy's to ~1e-13 relative, so the arg-min
    can only differ where two candidates are closer than that; fq_kl_search reports the relative gap to the
    runner-up and the Python layer flags gaps < 1e-9.  Here the gap is 8e-11 / 3e-11 and the choice flips with one
    ULP of one bin -- the kernel follows the oracle on both sides of the flip and flags both."""
    import warnings
    from quantization.mxnet_b200.quantize import distribution_calibrate as DC
    prev = ops.get_promotion()
    ops.set_promotion("nep50")          # the regime the histograms were tuned in (and the golden fixtures use)
    try:
        _near_tie_body(ops, DC, warnings)
    finally:
        ops.set_promotion(prev)


def _near_tie_body(ops, DC, warnings):
    want_best = []
    for h in _near_tie_histograms():
        div = O.kl_divergences(h, 256, 256, R.BINS, "nep50")
        d = np.where(np.isnan(div), np.inf, div)
        o = np.argsort(d, kind="stable")
        want_margin = (d[o[1]] - d[o[0]]) / abs(d[o[0]])
        assert sorted(o[:2].tolist()) == [1533, 1792] and want_margin < 1e-10
        margin = torch.empty(1, dtype=torch.float64, device="cuda")
        best, got = ops.kl_search(dev(h), 256, 256, R.BINS, promotion="nep50", margin=margin)
        assert abs(float(margin[0]) - want_margin) < 1e-12, (float(margin[0]), want_margin)
        assert int(best[0]) == int(o[0]) == O.kl_calibrate(h, 256, 256, R.BINS, "nep50")
        want_best.append(int(o[0]))
        with pytest.warns(RuntimeWarning, match="near-tie"):
            assert DC.kl_calibrate(h, 256, 256, R.BINS) in (1533, 1792)
    assert want_best == [1792, 1533]
    # an ordinary histogram has a comfortable margin and raises nothing
    h = R.kl_hist_cases()["relu"]
    margin = torch.empty(1, dtype=torch.float64, device="cuda")
    ops.kl_search(dev(h), 256, 256, R.BINS, promotion="nep50", margin=margin)
    assert float(margin[0]) > 1e-3
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert DC.kl_calibrate(h, 256, 256, R.BINS) == 1792


_ONLINE_MODE_SNIPPET = r"""
import sys
import numpy as np, torch
sys.path.insert(0, %r)
from oracle import fq_oracle as O
from quantization.mxnet_b200 import ops
r = np.random.RandomState(5)
for shape, bits, signed in [((128, 16, 32, 32), 8, False), ((128, 64, 8, 8), 4, True), ((32, 64), 8, False),
                            ((7, 36), 5, True), ((16, 3, 30, 30), 8, False), ((128, 32, 16, 16), 2, False),
                            ((64, 96, 28, 28), 8, False)]:
    x = r.standard_normal(shape).astype(np.float32)
    if not signed:
        x = np.abs(x)
    xd = torch.from_numpy(x).cuda()
    per = torch.empty(shape[0], device="cuda")
    for rep in range(3):            # the workspace must come back clean every time
        y, cur, qp = ops.forward_online(xd, bits, signed, ops.LO_NEG_MAX if signed else ops.LO_ZERO, per_sample=per)
        oy, _, ocur, oqp = O.fake_quant_input(x, bits, signed, None, "legacy", "conv")
        assert np.array_equal(y.cpu().numpy().view(np.uint32), oy.view(np.uint32)), (shape, rep)
        assert np.float32(cur.item()) == ocur and np.array_equal(qp.cpu().numpy(), np.array(oqp, np.float32))
        assert np.array_equal(per.cpu().numpy(), O.absmax_rows(x, shape[0]))
    # and the kernels that share the workspace still see it zeroed
    assert np.array_equal(ops.absmax_rows(xd, shape[0]).cpu().numpy(), O.absmax_rows(x, shape[0]))
print("ok")
"""


@pytest.mark.parametrize("mode", ["0", "2"])
def test_forward_online_every_execution_mode_is_bit_exact(mode):
    """FQ_ONLINE_MODE selects how a latency-bound tensor runs (csrc/fq_fused.cu): 0 = last-block finish + dependent
    streaming quantiser, 2 = dependent quantiser that finishes the range itself (default).  The choice is read once
    per process, hence the subprocess; both must match the oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FQ_ONLINE_MODE=mode)
    out = subprocess.run([sys.executable, "-c", _ONLINE_MODE_SNIPPET % root], capture_output=True, text=True,
                         timeout=600, env=env, cwd=root)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-3000:]


@pytest.mark.parametrize("shape,bits,signed", [((16, 8, 12, 12), 8, False), ((128, 16, 32, 32), 4, True),
                                               ((6, 37), 8, False), ((256, 64, 56, 56), 8, False)])
def test_forward_from_maxima_equals_the_online_path(ops, shape, bits, signed):
    """Data parallel online inputs: the quantiser is handed the all-gathered per-sample maxima and derives the Kahan
    mean and the scale itself -- same bits as the single-GPU online path (convert_conv2d.py:56-66) on the whole batch,
    on the one-launch path (small tensors) and on the three-launch path (large or ragged ones)."""
    x = rng(sum(shape)).standard_normal(shape).astype(F32)
    if not signed:
        x = np.abs(x)
    lo_mode = ops.LO_NEG_MAX if signed else ops.LO_ZERO
    maxima = dev(O.absmax_rows(x, shape[0]))
    y, cur, qp = ops.forward_from_maxima(dev(x), maxima, bits, signed, lo_mode)
    oy, _, ocur, oqp = O.fake_quant_input(x, bits, signed, None, "legacy", "conv")
    bits_equal(host(y), oy)
    assert F32(cur.item()) == ocur
    bits_equal(host(qp), np.array(oqp, F32))
    y2, cur2, qp2 = ops.forward_online(dev(x), bits, signed, lo_mode)
    bits_equal(host(y2), host(y))
