"""Property tests of the oracle (hypothesis): invariants the reference's arithmetic implies, checked on random
inputs so that a slip in the restatement shows up even where no golden vector happens to look."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st
from hypothesis.extra import numpy as hnp

from oracle import fq_oracle as O

F32 = np.float32
finite32 = st.floats(min_value=-1e6, max_value=1e6, allow_nan=False, allow_infinity=False, width=32)
vec = hnp.arrays(np.float32, st.integers(1, 300), elements=finite32)
CFG = dict(max_examples=60, deadline=None)


@settings(**CFG)
@given(vec)
def test_roundf_is_half_away_from_zero_and_odd(x):
    r = O.roundf(x)
    assert np.array_equal(r, np.trunc(r)) and np.all(np.abs(r.astype(np.float64) - x.astype(np.float64)) <= 0.5)
    assert np.array_equal(O.roundf(-x), -r)
    ties = np.abs(x - np.trunc(x)) == 0.5
    assert np.all(np.abs(r[ties]) == np.abs(np.trunc(x[ties])) + 1)


@settings(**CFG)
@given(vec, st.sampled_from([2, 3, 4, 8, 12, 16]), st.booleans(), st.sampled_from(["legacy", "nep50"]))
def test_fake_quant_input_codes_stay_in_range_and_are_monotone(x, bits, signed, promo):
    if not signed:
        x = np.abs(x)
    y, code, cur, (d, s, lo, hi) = O.fake_quant_input(x.reshape(1, -1), bits, signed, None, promo, "conv")
    q = O.qmax_of(bits, signed)
    assert np.array_equal(code, np.trunc(code))
    assert code.max() <= q and code.min() >= (-q if signed else 0)
    assert cur == F32(np.abs(x).max())                      # one sample: the mean of one maximum
    order = np.argsort(x.reshape(-1), kind="stable")
    assert np.all(np.diff(code.reshape(-1)[order]) >= 0)    # quantisation never reorders values
    assert np.array_equal(y, (code * s).astype(F32))
    if cur == 0:
        assert not np.any(y)                                # max_ == 0: divisor 1e-10, everything clips to 0


@settings(**CFG)
@given(hnp.arrays(np.float32, st.tuples(st.integers(1, 6), st.integers(1, 40)),
                  elements=st.floats(-100, 100, allow_nan=False, width=32)), st.sampled_from([2, 4, 8]))
def test_weight_quant_error_is_at_most_half_a_step_per_row(w, bits):
    y, code, s = O.fake_quant_weight(w, bits, "channel")
    q = O.qmax_of(bits, True)
    assert np.abs(code).max() <= q
    step = (s + F32(1e-10)).reshape(-1, 1).astype(np.float64)
    assert np.all(np.abs(y.astype(np.float64) - w.astype(np.float64)) <= 0.5 * step * (1 + 1e-5) + 1e-9 * q)
    # per-layer quantisation of a single row is the same thing
    y1, _, s1 = O.fake_quant_weight(w[:1], bits, "layer")
    assert np.array_equal(y1, y[:1]) and s1[0] == s[0]


@settings(**CFG)
@given(hnp.arrays(np.float32, st.integers(1, 2000), elements=st.floats(0, 50, allow_nan=False, width=32)),
       hnp.arrays(np.float32, st.integers(1, 2000), elements=st.floats(0, 50, allow_nan=False, width=32)),
       st.sampled_from(["legacy", "nep50"]), st.sampled_from([16, 100, 2048]))
def test_histogram_counts_add_up_and_concatenate(a, b, promo, bins):
    mx = F32(max(a.max(), 1e-3))                            # batch 0 freezes the max; batch 1 may exceed it
    ha = O.histogram_counts(a, bins, mx, promo)
    hb = O.histogram_counts(b, bins, mx, promo)
    hab = O.histogram_counts(np.concatenate([a, b]), bins, mx, promo)
    n = max(len(ha), len(hb), len(hab))
    pad = lambda h: np.pad(h, (0, n - len(h)))
    assert np.array_equal(pad(ha) + pad(hb), pad(hab))
    assert ha.sum() == np.count_nonzero(a) and hb.sum() == np.count_nonzero(b)       # zeros are dropped (:40)
    assert len(hab) <= bins + 1 and hab.min() >= 0


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.sampled_from([32, 64]), st.sampled_from(["legacy", "nep50"]))
def test_kl_search_is_invariant_to_power_of_two_scaling(seed, levels, promo):
    r = np.random.RandomState(seed)
    bins = 256
    h = np.floor(r.gamma(2.0, 200.0, bins) * np.exp(-np.arange(bins) / 60.0)).astype(F32)
    h[0] += 1
    best = O.kl_calibrate(h, levels, levels, bins, promo)
    assert levels <= best < bins
    d1 = O.kl_divergences(h, levels, levels, bins, promo)
    d4 = O.kl_divergences((h * F32(4)).astype(F32), levels, levels, bins, promo)
    # every sum scales exactly by 4 and every ratio is unchanged: bit-identical divergences
    assert np.array_equal(np.nan_to_num(d1, nan=-1.0), np.nan_to_num(d4, nan=-1.0))
    assert O.kl_calibrate((h * F32(4)).astype(F32), levels, levels, bins, promo) == best


@settings(**CFG)
@given(hnp.arrays(np.float32, st.integers(1, 200), elements=st.floats(0, 1e4, allow_nan=False, width=32)))
def test_kahan_mean_properties(v):
    m = O.mean_kahan_f32(v)
    assert v.min() <= m <= v.max() or np.isclose(m, v.mean(), rtol=1e-6)
    exact = np.float64(v.astype(np.float64).sum()) / len(v)
    assert abs(np.float64(m) - exact) <= 2.0 ** -22 * max(abs(exact), 1e-30)       # compensated: ~1 ulp of the mean
    c = np.full(len(v), v[0], F32)
    assert abs(np.float64(O.mean_kahan_f32(c)) - np.float64(v[0])) <= 2.0 ** -23 * abs(np.float64(v[0]))


@settings(max_examples=30, deadline=None)
@given(hnp.arrays(np.float32, st.tuples(st.integers(1, 4), st.integers(1, 3), st.just(3), st.just(3)),
                  elements=st.floats(-4, 4, allow_nan=False, width=32)), st.sampled_from(["F23", "F43", "F63"]))
def test_winograd_transform_scales_exactly_and_backward_is_near_identity(w, name):
    # "x2 is exact" only holds while no product or partial sum is subnormal (3.788e-42 * G loses bits that the
    # doubled input keeps), so tiny magnitudes are flushed to zero before the property is checked
    w = np.where(np.abs(w) < F32(2.0 ** -60), F32(0), w).astype(F32)
    G, GI, GTI = O.winograd_matrices(name)
    U = O.wino_transform(w, G)
    assert np.array_equal(O.wino_transform((w * F32(2)).astype(F32), G), (U * F32(2)).astype(F32))
    assert np.array_equal(O.wino_transform(-w, G), -U)
    g = O.wino_backward(w, name)
    assert np.abs(g - w).max() <= 1e-5 * max(1.0, np.abs(w).max())


@settings(**CFG)
@given(st.floats(0, 1e4, width=32), st.floats(0, 1e4, width=32), st.sampled_from(["legacy", "nep50"]))
def test_ema_stays_between_state_and_current(state, cur, promo):
    out = O.ema_scalar(np.array([state], F32), np.array([cur], F32), 0.9, promo)[0]
    lo, hi = min(state, cur), max(state, cur)
    assert lo * (1 - 1e-6) - 1e-30 <= out <= hi * (1 + 1e-6) + 1e-30
