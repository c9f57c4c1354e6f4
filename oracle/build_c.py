"""Compile oracle/fq_oracle.c (the C restatement; test infrastructure) into oracle/build/libfq_oracle.so.

The reference is pure Python, so there is nothing of its own to compile into oracle/_ref/; what runs
"verbatim" is quantize/distribution_calibrate.py, executed by oracle/make_golden.py in the build container.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fq_oracle.c")
LIB = os.path.join(HERE, "build", "libfq_oracle.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                           "-fvisibility=hidden", SRC, "-o", LIB, "-lm"])
    return LIB


_lib = None
_f = ctypes.POINTER(ctypes.c_float)
_d = ctypes.POINTER(ctypes.c_double)
_l = ctypes.POINTER(ctypes.c_int64)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.fqo_absmax_rows.argtypes = [_f, ctypes.c_int64, ctypes.c_int64, _f]
        L.fqo_mean_kahan.argtypes = [_f, ctypes.c_int64]
        L.fqo_mean_kahan.restype = ctypes.c_float
        L.fqo_fake_quant_scalar.argtypes = [_f, ctypes.c_int64, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                            ctypes.c_float, ctypes.c_int, _f, _f]
        L.fqo_fake_quant_rows.argtypes = [_f, ctypes.c_int64, ctypes.c_int64, _f, _f, _f]
        L.fqo_hist_counts.argtypes = [_f, ctypes.c_int64, ctypes.c_float, ctypes.c_float, ctypes.c_int, _l]
        L.fqo_kl_calibrate.argtypes = [_f, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _d]
        L.fqo_kl_calibrate.restype = ctypes.c_int
        L.fqo_wino_weight.argtypes = [_f, ctypes.c_int64, ctypes.c_int64, _f, _f, _f, ctypes.c_int, ctypes.c_int, _f, _f]
        L.fqo_channel_stats.argtypes = [_f, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _f, _f]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def absmax_rows(x, rows):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(rows, np.float32)
    lib().fqo_absmax_rows(_p(x, _f), rows, x.size // rows, _p(out, _f))
    return out


def mean_kahan(v):
    v = np.ascontiguousarray(v, np.float32)
    return np.float32(lib().fqo_mean_kahan(_p(v, _f), v.size))


def fake_quant_scalar(x, d, s, lo=None, hi=None):
    x = np.ascontiguousarray(x, np.float32)
    y, c = np.empty_like(x), np.empty_like(x)
    clip = hi is not None
    lib().fqo_fake_quant_scalar(_p(x, _f), x.size, d, s, lo if clip else 0.0, hi if clip else 0.0, int(clip),
                                _p(y, _f), _p(c, _f))
    return y, c


def fake_quant_rows(x, rows, scale):
    x = np.ascontiguousarray(x, np.float32)
    scale = np.ascontiguousarray(scale, np.float32).reshape(-1)
    y, c = np.empty_like(x), np.empty_like(x)
    lib().fqo_fake_quant_rows(_p(x, _f), rows, x.size // rows, _p(scale, _f), _p(y, _f), _p(c, _f))
    return y, c


def histogram_counts(x, bins, max_, sc):
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    counts = np.zeros(bins + 1, np.int64)
    lib().fqo_hist_counts(_p(x, _f), x.size, np.float32(max_), np.float32(sc), bins, _p(counts, _l))
    return counts


def kl_calibrate(data, levels, min_bins, bins, promotion="nep50"):
    data = np.ascontiguousarray(data, np.float32)
    div = np.full(bins, np.nan, np.float64)
    best = lib().fqo_kl_calibrate(_p(data, _f), data.size, levels, min_bins, bins, int(promotion == "nep50"), _p(div, _d))
    return best, div


def wino_weight(w, G, GI, GTI, bits):
    w = np.ascontiguousarray(w, np.float32)
    G, GI, GTI = (np.ascontiguousarray(m, np.float32) for m in (G, GI, GTI))
    wq, s = np.empty_like(w), np.empty(w.shape[0], np.float32)
    lib().fqo_wino_weight(_p(w, _f), w.shape[0], w.shape[1], _p(G, _f), _p(GI, _f), _p(GTI, _f), G.shape[0], bits,
                          _p(wq, _f), _p(s, _f))
    return wq, s


def channel_stats(y):
    y = np.ascontiguousarray(y, np.float32)
    n, c = y.shape[0], y.shape[1]
    mean, var = np.empty(c, np.float32), np.empty(c, np.float32)
    lib().fqo_channel_stats(_p(y, _f), n, c, y.size // (n * c), _p(mean, _f), _p(var, _f))
    return mean, var


if __name__ == "__main__":
    print(build(force=True))
