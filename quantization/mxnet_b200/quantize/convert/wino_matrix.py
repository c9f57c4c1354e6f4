"""Kernel-transform matrices G of Winograd F(2x2,3x3), F(4x4,3x3) and F(6x6,3x3) (the reference's
``quantize/convert/wino_matrix.py:28-60``), as float32 like ``nd.array`` makes them, plus the float32
pseudo-inverses the Winograd-domain weight quantiser multiplies back with (``convert_conv2d.py:80-82``:
``np.linalg.pinv`` of the float32 matrices, computed on the host)."""
from fractions import Fraction as _F

import numpy as np

__all__ = ['Winograd_G', 'winograd_matrices']


def _rows(*rows):
    return np.array([[float(_F(v)) for v in r] for r in rows], dtype=np.float32)


Winograd_G = {
    # F(m x m, 3 x 3): (m + 2) x 3
    "F23": _rows(("1", "0", "0"), ("1/2", "1/2", "1/2"), ("1/2", "-1/2", "1/2"), ("0", "0", "1")),
    "F43": _rows(("1/4", "0", "0"), ("-1/6", "-1/6", "-1/6"), ("-1/6", "1/6", "-1/6"), ("1/24", "1/12", "1/6"),
                 ("1/24", "-1/12", "1/6"), ("0", "0", "1")),
    "F63": _rows(("1", "0", "0"), ("-2/9", "-2/9", "-2/9"), ("-2/9", "2/9", "-2/9"), ("1/90", "1/45", "2/45"),
                 ("1/90", "-1/45", "2/45"), ("32/45", "16/45", "8/45"), ("32/45", "-16/45", "8/45"), ("0", "0", "1")),
}


def winograd_matrices(name):
    """(G [a,3], GI = pinv(G) [3,a], GTI = pinv(G.T) [a,3]) as float32 NumPy arrays."""
    G = Winograd_G[name]
    GI = np.linalg.pinv(G).astype(np.float32)
    GTI = np.linalg.pinv(np.ascontiguousarray(G.T)).astype(np.float32)
    return G, np.ascontiguousarray(GI), np.ascontiguousarray(GTI)
