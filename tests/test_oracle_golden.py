"""The oracle restatement vs. outputs of the reference ITSELF (tests/golden/*.npz,
written by oracle/make_golden.py from /root/reference/quantize/distribution_calibrate.py)."""
import os

import numpy as np
import pytest

from oracle import fq_oracle as O
from oracle import golden_recipes as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HIST = np.load(os.path.join(GOLD, "hist_nep50.npz"))
KL = np.load(os.path.join(GOLD, "kl_nep50.npz"))


def test_fixture_regime():
    assert str(HIST["regime"]) == "nep50" and str(KL["regime"]) == "nep50"


@pytest.mark.parametrize("name", sorted(R.hist_cases().keys()))
def test_histogram_matches_reference(name):
    batches = R.hist_cases()[name]
    fm_max = None
    acc = 0
    err = ""
    for b, fm in enumerate(batches):
        h, m = O.discrete_histogram(fm, R.BINS, fm_max, promotion="nep50")
        want = HIST["hist/%s/batch%d" % (name, b)]
        assert h.dtype == np.float32 and h.shape == want.shape
        assert np.array_equal(h, want)
        if fm_max is None:
            fm_max = m
        try:
            acc = acc + h
        except ValueError as e:
            err = type(e).__name__
            break
    assert np.float32(fm_max) == HIST["hist/%s/max" % name]
    assert err == str(HIST["hist/%s/error" % name])
    if not err:
        assert np.array_equal(acc, HIST["hist/%s/acc" % name])


def test_accumulate_histograms_helper():
    name = "three_batches_frozen_max"
    acc, mx = O.accumulate_histograms(R.hist_cases()[name], R.BINS)
    assert np.array_equal(acc, HIST["hist/%s/acc" % name]) and mx == HIST["hist/%s/max" % name]


_KL_CASES = [(n, l) for n, ls in R.KL_LEVELS.items() for l in ls]


@pytest.mark.parametrize("name,levels", _KL_CASES)
def test_kl_best_bin_matches_reference(name, levels):
    h = KL["kl/%s/hist" % name]
    assert np.array_equal(h, R.kl_hist_cases()[name])       # recipes are deterministic
    best = O.kl_calibrate(h, levels, levels, R.BINS, promotion="nep50")
    assert best == int(KL["kl/%s/L%d/best" % (name, levels)])


@pytest.mark.parametrize("name", sorted(R.KL_LEVELS.keys()))
def test_kl_windows_match_reference(name):
    h = KL["kl/%s/hist" % name]
    levels = R.KL_LEVELS[name][0]
    div = O.kl_divergences(h, levels, levels, R.BINS, promotion="nep50")
    for lo, hi in R.KL_WINDOWS:
        best, bd = max(lo, levels), np.inf
        for i in range(max(lo, levels), hi):
            if div[i] < bd:
                bd, best = div[i], i
        assert best == int(KL["kl/%s/L%d/win_%d_%d" % (name, levels, lo, hi)])
