"""The CUDA kernels replace roundf(v / d) by rintf(v * RN(1/d)) guarded by a distance-to-tie test
(csrc/fq_common.cuh QDiv).  This replays that arithmetic in NumPy float32 and checks it against the
oracle's IEEE divide + roundf on tie-heavy and random data: the guarded fast path must never
disagree, and must be taken almost always."""
import numpy as np
import pytest

from oracle import fq_oracle as O

F32 = np.float32


def qdiv_code(v, d):
    v = v.astype(F32)
    d = F32(d)
    r = F32(1) / d
    q0 = (v * r).astype(F32)
    t = np.rint(q0).astype(F32)
    with np.errstate(invalid="ignore"):
        e = np.abs((q0 - t).astype(F32))
        fast = (F32(0.5) - e).astype(F32) > (np.abs(q0) * F32(2.0 ** -21)).astype(F32)
    slow = O.roundf((v / d).astype(F32))
    return np.where(fast, t, slow).astype(F32), fast


@pytest.mark.parametrize("bits", [2, 4, 8, 12, 16])
def test_guarded_reciprocal_equals_ieee_divide_then_roundf(bits):
    r = np.random.RandomState(bits)
    qmax = 2 ** bits - 1
    unguarded_wrong = 0
    for trial in range(60):
        max_ = F32(abs(r.standard_normal()) * 10 ** r.uniform(-6, 4))
        d, s, lo, hi = O.input_qparams(max_, bits, False, "legacy" if trial % 2 else "nep50")
        k = r.randint(0, qmax + 1, 200_000).astype(F32)
        v = ((k + F32(0.5)) * d).astype(F32)                 # on and around every rounding tie
        v[::3] = np.nextafter(v[::3], F32(np.inf))
        v[1::3] = np.nextafter(v[1::3], F32(-np.inf))
        v[::4] = (r.uniform(0, 1, len(v[::4])) * max_).astype(F32)
        v = O.clip(v, lo, hi)
        got, fast = qdiv_code(v, d)
        want = O.roundf((v / d).astype(F32))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        unguarded = np.rint((v * (F32(1) / d)).astype(F32))
        unguarded_wrong += int((unguarded != want).sum())
    # without the guard the reciprocal path does get ties wrong: the guard is doing real work
    assert unguarded_wrong > 0


def test_fast_path_is_taken_almost_always_on_ordinary_data():
    r = np.random.RandomState(0)
    v = np.abs(r.standard_normal(2_000_000)).astype(F32)
    d, s, lo, hi = O.input_qparams(F32(3.0), 8, False, "legacy")
    got, fast = qdiv_code(O.clip(v, lo, hi), d)
    want = O.roundf((O.clip(v, lo, hi) / d).astype(F32))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert fast.mean() > 0.9995


def test_markstein_correction_gives_the_correctly_rounded_double_quotient():
    """csrc/fq_calib.cu DDiv: y = RN(1/b), q0 = RN(a*y), r = fma(-b, q0, a), q = fma(r, y, q0) must equal a / b.
    FMA is emulated with exact rational arithmetic (float(Fraction) rounds to nearest even)."""
    import random
    import struct
    from fractions import Fraction as Fr

    def fma(x, y, z):
        return float(Fr(x) * Fr(y) + Fr(z))

    def ddiv(a, b):
        y = 1.0 / b
        q0 = a * y
        return fma(fma(-b, q0, a), y, q0)
    rnd = random.Random(3)
    for levels in (256, 128, 16, 255):                    # t = j * levels / i
        for i in list(range(2, 120)) + rnd.sample(range(120, 2048), 60):
            for j in rnd.sample(range(i), min(i, 25)):
                assert ddiv(float(j * levels), float(i)) == float(j * levels) / float(i)
    for _ in range(40000):                                # Q_j / sum(Q)
        b = rnd.uniform(1, 2) * 2.0 ** rnd.randint(-5, 40)
        a = b * rnd.random() * 2.0 ** -rnd.randint(0, 30)
        assert ddiv(a, b) == a / b
    for e in range(-2, 3):                                # divisors whose significand is all ones
        b = struct.unpack("d", struct.pack("Q", (0x3ff + e) << 52 | 0xfffffffffffff))[0]
        for _ in range(1500):
            a = rnd.uniform(0, 4) * b
            assert ddiv(a, b) == a / b
