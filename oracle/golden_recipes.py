"""Deterministic input recipes shared by oracle/make_golden.py and tests/.

TEST INFRASTRUCTURE ONLY.  Inputs are regenerated from these recipes (legacy
``numpy.random.RandomState`` streams are frozen across NumPy versions), so
``tests/golden/*.npz`` only has to store the reference's OUTPUTS.
"""
import numpy as np

F32 = np.float32
BINS = 2048


def relu_normal(seed, n, scale=1.0):
    r = np.random.RandomState(seed)
    return np.maximum(r.standard_normal(n), 0).astype(F32) * F32(scale)


def relu6_like(seed, n):
    r = np.random.RandomState(seed)
    return np.clip(r.standard_normal(n) * 3.0, 0, 6).astype(F32)


# name -> (callable producing the list of batches, explicit first max or None)
def hist_cases():
    cases = {}
    cases["relu_50k"] = [relu_normal(1, 50_000)]
    cases["relu6_80k"] = [relu6_like(2, 80_000)]
    cases["three_batches_frozen_max"] = [relu_normal(3, 40_000), relu_normal(4, 40_000, 1.5),
                                         relu_normal(5, 40_000, 0.5)]
    # max >= 256: 1e-5 vanishes in fp32, the element equal to max can land in bin 2048
    for k, sc in enumerate((100.0, 173.0, 300.0, 517.0, 1000.0, 4099.0)):
        cases["bigmax_%d" % k] = [relu_normal(10 + k, 30_000, sc)]
    cases["tiny_values"] = [relu_normal(20, 20_000, 1e-6)]
    cases["denormal_values"] = [relu_normal(21, 5_000, 1e-41)]
    cases["all_equal"] = [np.full(10_000, 0.75, dtype=F32)]
    cases["mostly_zero"] = [np.where(np.arange(20_000) % 97 == 0, 2.5, 0).astype(F32)]
    # > 2^24 in one bin after accumulation: fp32 running adds round
    big = np.full(9_000_001, 1.0, dtype=F32)
    big[:7] = [0.5, 0.25, 0.125, 2.0, 1.5, 0.0, 1.75]
    cases["counts_over_2p24"] = [big, big[:8_999_999], big[:9_000_000]]
    cases["single_element"] = [np.array([3.0], dtype=F32)]
    return cases


def kl_hist_cases():
    """name -> float32 histogram (length 2048 or 2049)."""
    from . import fq_oracle as O   # only used to build inputs, not outputs
    h = {}
    h["relu"] = O.discrete_histogram(relu_normal(31, 400_000), BINS)[0]
    h["relu6"] = O.discrete_histogram(relu6_like(32, 400_000), BINS)[0]
    h["flat"] = np.full(BINS, 100.0, dtype=F32)
    spike = np.zeros(BINS, dtype=F32)
    spike[300] = 1000.0
    h["spike"] = spike
    r = np.random.RandomState(33)
    h["exp_decay"] = np.floor(1e6 * np.exp(-np.arange(BINS) / 150.0) * (0.5 + r.rand(BINS))).astype(F32)
    h["sparse"] = (r.rand(BINS) < 0.05).astype(F32) * np.floor(r.rand(BINS) * 50).astype(F32)
    # totals far above 2^24: sequential fp32 sums round, order matters
    h["huge_counts"] = (np.floor(3e7 * np.exp(-np.arange(BINS) / 400.0)) + r.randint(0, 1000, BINS)).astype(F32)
    h["len2049"] = np.concatenate([h["relu6"], np.array([37.0], dtype=F32)])
    h["all_zero"] = np.zeros(BINS, dtype=F32)
    h["non_integer"] = (r.rand(BINS) * 10).astype(F32)
    # long-tailed activations: the KL optimum is an interior threshold
    x = relu_normal(34, 600_000)
    x[::5000] *= 9.0
    h["outliers"] = O.discrete_histogram(x, BINS)[0]
    r2 = np.random.RandomState(35)
    h["lognormal"] = O.discrete_histogram(np.exp(r2.standard_normal(500_000) * 1.2).astype(F32), BINS)[0]
    xb = np.concatenate([relu_normal(36, 300_000, 0.2), (3.0 + 0.1 * r2.standard_normal(50_000)).astype(F32),
                         np.array([40.0], dtype=F32)])
    h["bimodal_outlier"] = O.discrete_histogram(xb, BINS)[0]
    return h


KL_LEVELS = {
    "relu": (256, 128, 16),
    "relu6": (256, 64),
    "flat": (256, 4),
    "spike": (256, 128),
    "exp_decay": (256, 32),
    "sparse": (128,),
    "huge_counts": (256, 128, 8),
    "len2049": (256,),
    "all_zero": (256,),
    "non_integer": (256,),
    "outliers": (256, 128, 16),
    "lognormal": (256, 128, 4),
    "bimodal_outlier": (256, 32),
}
# extra (min_bins, bins) windows: arg-min restricted to sub-ranges pins more of the curve
KL_WINDOWS = ((300, 700), (700, 1200), (1200, 2048))
