"""The one-pass float64 shifted-moment formula of csrc/fq_stats.cu, replayed in NumPy on the CPU, against the
oracle's restatement of convert_conv2d.py:150-153 (two-pass sequential Kahan fp32 sums).  This pins the claimed
bound (mean <= 1 ULP -- of the mean magnitude when the channel's values cancel --, var <= 2 ULP) without a GPU, including the data-parallel combination of per-rank records."""
import numpy as np
import pytest

from oracle import build_c as C
from oracle import fq_oracle as O

F32, F64 = np.float32, np.float64


def records(y):
    """{n, S1, S2, K} per channel of one rank's shard, as channel_stats_kernel accumulates them."""
    c = y.shape[1]
    f = np.moveaxis(y.reshape(y.shape[0], c, -1), 1, 0).reshape(c, -1).astype(F64)
    K = f[:, 0].copy()
    d = f - K[:, None]
    return np.stack([np.full(c, f.shape[1], F64), d.sum(axis=1), (d * d).sum(axis=1), K], axis=1)


def combine(recs):
    """stats_combine() of csrc/fq_stats.cu over a list of per-rank [C, 4] records."""
    n = sum(r[:, 0] for r in recs)
    sy = sum(r[:, 1] + r[:, 0] * r[:, 3] for r in recs)
    nf = n.astype(F32)
    mean = (sy.astype(F32) / nf).astype(F32)
    m2 = 0.0
    for r in recs:
        d = mean.astype(F64) - r[:, 3]
        m2 = m2 + (r[:, 2] + d * (r[:, 0] * d - 2.0 * r[:, 1]))
    return mean, (np.maximum(m2, 0.0).astype(F32) / nf).astype(F32)


def ulp(a, b):
    return int(np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64)).max())


def mean_close(got, want, y):
    """Kahan's bound is relative to sum |y|, not to |sum y|: a channel whose positive and negative values cancel
    has a mean of few ULP(mean |y|), not few ULP of itself.  1 ULP of the result + 2^-23 of the mean magnitude."""
    mag = np.abs(y.astype(F64)).mean(axis=(0, 2, 3))
    tol = np.spacing(np.abs(want)).astype(F64) + 2.0 ** -23 * mag
    return bool(np.all(np.abs(got.astype(F64) - want.astype(F64)) <= tol))


def test_c_oracle_equals_numpy_oracle():
    r = np.random.RandomState(3)
    for shape in [(4, 3, 5, 5), (7, 6, 3, 9), (2, 1, 1, 1), (9, 4, 7, 7)]:
        y = (r.standard_normal(shape) * 3 + 1).astype(F32)
        m, v = O.channel_stats(y)
        mc, vc = C.channel_stats(y)
        assert np.array_equal(m, mc) and np.array_equal(v, vc)


@pytest.mark.parametrize("ranks", [1, 2, 4, 8])
@pytest.mark.parametrize("offset,scale", [(0.0, 1.0), (0.5, 2.0), (300.0, 0.5), (-1e4, 3.0), (1e-3, 1e-4)])
def test_shifted_moments_within_two_ulp_of_the_kahan_restatement(ranks, offset, scale):
    r = np.random.RandomState(ranks * 100 + int(abs(offset)) % 97)
    y = (r.standard_normal((8 * ranks, 12, 14, 14)) * scale + offset).astype(F32)
    want_m, want_v = C.channel_stats(y)
    got_m, got_v = combine([records(s) for s in np.array_split(y, ranks, axis=0)])
    assert mean_close(got_m, want_m, y)
    if offset != 0.0:
        assert ulp(got_m, want_m) <= 1          # no cancellation: 1 ULP of the mean itself
    assert ulp(got_v, want_v) <= 2


def test_constant_and_tiny_channels():
    y = np.full((4, 3, 6, 6), 777.125, F32)
    y[:, 1] = 0.0
    y[:, 2] = 1e-30
    m, v = combine([records(y)])
    wm, wv = C.channel_stats(y)
    assert np.array_equal(m, wm) and np.array_equal(v, wv)
    # uneven shards (last batch smaller) still combine exactly like the whole batch
    r = np.random.RandomState(11)
    y = (r.standard_normal((10, 5, 7, 7)) + 2).astype(F32)
    m, v = combine([records(y[:7]), records(y[7:])])
    wm, wv = C.channel_stats(y)
    assert ulp(m, wm) <= 1 and ulp(v, wv) <= 2
