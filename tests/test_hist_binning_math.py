"""The histogram kernel bins with a branch-free sequence (csrc/fq_calib.cu, HistBins::add):

    p = v > 0 ; t = min(v, max_) ; q = cvt.rzi.u32(t * sc) ; q = min(q, bins) ; count[p ? q : trash] += 1

instead of the reference's clip(0, max_) -> drop zeros -> trunc(v * sc) (distribution_calibrate.py:39-45).  This replays
the sequence in NumPy float32 and checks it against the oracle on the values where the two formulations could
part: zeros of both signs, negatives, denormals, values at and above the frozen max, bin boundaries."""
import numpy as np
import pytest

from oracle import fq_oracle as O

F32 = np.float32


def kernel_bins(x, bins, max_, promotion):
    x = np.asarray(x, F32)
    sc = O.hist_scale(F32(max_), bins, promotion)
    with np.errstate(invalid="ignore", over="ignore"):
        p = x > 0                                                  # setp.gt.f32: false for NaN, -0.0, negatives
        t = np.where(np.isnan(x), F32(max_), np.minimum(x, F32(max_))).astype(F32)      # PTX min.f32(NaN, b) = b
        prod = (t * sc).astype(F32)
        # cvt.rzi.u32.f32 saturates: negatives and NaN -> 0, huge -> 2^32 - 1
        q = np.where(np.isnan(prod) | (prod <= 0), 0, np.minimum(np.trunc(prod.astype(np.float64)), 2.0 ** 32 - 1)).astype(np.uint64)
        q = np.minimum(q, bins)
    return np.bincount(q[p].astype(np.int64), minlength=bins + 1)[: bins + 1]


def edge_values(max_, bins, seed):
    r = np.random.RandomState(seed)
    step = np.float64(max_) / bins
    k = r.randint(0, bins + 1, 4000)
    on = (k * step).astype(F32)                                    # on and around every bin boundary
    vals = np.concatenate([
        on, np.nextafter(on, F32(np.inf)), np.nextafter(on, F32(-np.inf)),
        np.array([0.0, -0.0, -1.0, -1e-30, 1e-45, 1e-38, max_, np.nextafter(F32(max_), F32(np.inf)), 2 * max_, 1e30],
                 F32),
        (r.standard_normal(4000) * max_).astype(F32),
    ]).astype(F32)
    return vals


@pytest.mark.parametrize("promotion", ["legacy", "nep50"])
@pytest.mark.parametrize("bins", [1, 100, 2048])
@pytest.mark.parametrize("max_", [1e-3, 0.75, 6.0, 255.9999, 256.0, 1000.5, 3e7])
def test_branch_free_binning_equals_clip_drop_trunc(promotion, bins, max_):
    max_ = F32(max_)
    x = edge_values(max_, bins, int(max_ * 7) % 1000 + bins)
    want = O.histogram_counts(np.maximum(x, 0), bins, max_, promotion)      # the reference asserts min >= 0 (:35)
    got = kernel_bins(x, bins, max_, promotion)
    assert got[len(want):].sum() == 0
    assert np.array_equal(got[:len(want)], want)
    assert got.sum() == np.count_nonzero(np.maximum(x, 0))


def test_products_never_pass_the_last_counter():
    """0 < v <= max_ keeps trunc(v * sc) <= bins, so the clamp in the kernel is for memory safety only."""
    r = np.random.RandomState(1)
    for promotion in ("legacy", "nep50"):
        for bins in (1, 7, 2048, 8192):
            maxes = np.abs(r.standard_normal(20000) * 10.0 ** r.uniform(-6, 7, 20000)).astype(F32)
            maxes = maxes[maxes > 0]
            sc = np.array([O.hist_scale(m, bins, promotion) for m in maxes[:3000]], F32)
            top = (maxes[:3000] * sc).astype(F32)
            assert np.all(np.trunc(top) <= bins)
