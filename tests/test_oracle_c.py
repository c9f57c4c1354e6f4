"""The C (+OpenMP) restatement vs. the NumPy oracle and the reference's golden vectors."""
import os

import numpy as np
import pytest

from oracle import build_c as C
from oracle import fq_oracle as O
from oracle import golden_recipes as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F32 = np.float32


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a, F32).view(np.uint32), np.ascontiguousarray(b, F32).view(np.uint32))


def test_c_range_and_fake_quant_match_numpy_oracle():
    r = np.random.RandomState(0)
    x = (r.standard_normal((64, 3, 9, 9)) * 2).astype(F32)
    assert np.array_equal(C.absmax_rows(x, 64), O.absmax_rows(x, 64))
    per = O.absmax_rows(x, 64)
    assert C.mean_kahan(per) == O.mean_kahan_f32(per)
    for bits, signed in ((8, False), (4, True), (16, False), (2, True)):
        d, s, lo, hi = O.input_qparams(F32(2.5), bits, signed, "legacy")
        k = r.randint(-200, 200, x.size).astype(F32)
        t = ((k + F32(0.5)) * d).astype(F32).reshape(x.shape)     # rounding ties
        for data in (x, t):
            y, c = C.fake_quant_scalar(data, d, s, lo, hi)
            wy, wc = O.fake_quant_scalar(data, d, s, lo, hi)
            assert same_bits(y, wy) and same_bits(c, wc)
    w = (r.standard_normal((32, 27)) * 0.1).astype(F32)
    sc, dd, _ = O.weight_scales(w, 32, 8)
    y, c = C.fake_quant_rows(w, 32, sc)
    wy, wc = O.fake_quant_rows(w, 32, sc, dd)
    assert same_bits(y, wy) and same_bits(c, wc)


@pytest.mark.parametrize("name", ["relu_50k", "bigmax_2", "three_batches_frozen_max", "tiny_values"])
def test_c_histogram_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, "hist_nep50.npz"))
    batches = R.hist_cases()[name]
    mx = g["hist/%s/max" % name]
    for b, fm in enumerate(batches):
        want = g["hist/%s/batch%d" % (name, b)]
        got = C.histogram_counts(fm, R.BINS, mx, O.hist_scale(mx, R.BINS, "nep50"))
        assert np.array_equal(got[:len(want)], want.astype(np.int64)) and got[len(want):].sum() == 0


@pytest.mark.parametrize("name,levels", [("relu", 256), ("outliers", 16), ("lognormal", 128), ("huge_counts", 256),
                                         ("len2049", 256), ("all_zero", 256), ("non_integer", 256)])
def test_c_kl_matches_reference_golden(name, levels):
    g = np.load(os.path.join(GOLD, "kl_nep50.npz"))
    h = g["kl/%s/hist" % name]
    best, div = C.kl_calibrate(h, levels, levels, R.BINS, "nep50")
    assert best == int(g["kl/%s/L%d/best" % (name, levels)])
    want = O.kl_divergences(h, levels, levels, R.BINS, "nep50")
    assert np.allclose(div[levels:], want[levels:], rtol=1e-12, atol=1e-14, equal_nan=True)
    best_l, _ = C.kl_calibrate(h, levels, levels, R.BINS, "legacy")
    assert best_l == O.kl_calibrate(h, levels, levels, R.BINS, "legacy")


@pytest.mark.parametrize("name", ["F23", "F43", "F63"])
def test_c_winograd_weight_path_with_real_fmaf_matches_numpy_oracle(name):
    """The C restatement uses the hardware's single-rounding fmaf; the NumPy oracle emulates it in extended
    precision.  Bit-equal results on ~10^6 fused operations pin the emulation."""
    r = np.random.RandomState(9)
    w = (r.standard_normal((24, 17, 3, 3)) * r.choice([1e-3, 0.1, 3.0], (24, 1, 1, 1))).astype(np.float32)
    w[1] = 0
    G, GI, GTI = O.winograd_matrices(name)
    for bits in (8, 4):
        want, want_s, _, _ = O.fake_quant_weight_wino(w, name, bits)
        got, got_s = C.wino_weight(w, G, GI, GTI, bits)
        assert np.array_equal(got_s.view(np.uint32), want_s.view(np.uint32))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
