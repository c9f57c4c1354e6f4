"""Winograd-domain weight quantisation: the oracle restatement (convert_conv2d.py:71-83) against float64
linear algebra, and the host-side matrices against the reference's literals (wino_matrix.py:28-60)."""
import numpy as np
import pytest

from oracle import fq_oracle as O

F32 = np.float32


@pytest.mark.parametrize("name,a", [("F23", 4), ("F43", 6), ("F63", 8)])
def test_matrices_and_pseudo_inverses(name, a):
    from quantization.mxnet_b200.quantize import convert
    from quantization.mxnet_b200.quantize.convert import wino_matrix
    assert convert.wino_matrix is wino_matrix                      # exported like the reference's submodule
    G, GI, GTI = wino_matrix.winograd_matrices(name)
    oG, oGI, oGTI = O.winograd_matrices(name)
    assert G.shape == (a, 3) and GI.shape == (3, a) and GTI.shape == (a, 3)
    assert G.dtype == GI.dtype == GTI.dtype == F32
    for got, want in ((G, oG), (GI, oGI), (GTI, oGTI)):
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(wino_matrix.Winograd_G[name], G)
    assert np.allclose(GI.astype(np.float64) @ G.astype(np.float64), np.eye(3), atol=1e-6)
    assert np.allclose(G.T.astype(np.float64) @ GTI.astype(np.float64), np.eye(3), atol=1e-6)


@pytest.mark.parametrize("name", ["F23", "F43", "F63"])
def test_oracle_follows_the_reference_formula(name):
    r = np.random.RandomState(3)
    w = (r.standard_normal((6, 5, 3, 3)) * 0.2).astype(F32)
    G, GI, GTI = (m.astype(np.float64) for m in O.winograd_matrices(name))
    wq, s, U, Uq = O.fake_quant_weight_wino(w, name, 8)
    # nd.dot(nd.dot(G, w.transpose(2,3,0,1)).transpose(2,3,0,1), G.T) in float64
    U64 = np.einsum("pr,oirc,qc->oipq", G, w.astype(np.float64), G)
    assert np.abs(U - U64).max() <= 4e-7 * np.abs(U64).max()
    m = np.abs(U).reshape(6, -1).max(axis=1)
    assert np.array_equal(s, (m / F32(127)).astype(F32))
    codes = Uq / s.reshape(6, 1, 1, 1)
    assert np.abs(codes - np.round(codes)).max() < 1e-4 and np.abs(codes).max() <= 127.001
    wq64 = np.einsum("rp,oipq,qc->oirc", GI, Uq.astype(np.float64), GTI)
    assert np.abs(wq - wq64).max() <= 1e-6 * max(1.0, np.abs(wq64).max())
    # quantisation error is bounded by half a step per Winograd-domain element, pushed through the inverse
    bound = 0.5 * s.max() * np.abs(GI).sum(axis=1).max() * np.abs(GTI).sum(axis=0).max()
    assert np.abs(wq - w).max() <= bound * 1.01 + 1e-6


@pytest.mark.parametrize("name", ["F23", "F43", "F63"])
def test_oracle_backward_is_the_adjoint_of_the_four_products(name):
    r = np.random.RandomState(5)
    g = r.standard_normal((4, 3, 3, 3)).astype(F32)
    G, GI, GTI = (m.astype(np.float64) for m in O.winograd_matrices(name))
    dw = O.wino_backward(g, name)
    dU = np.einsum("rp,oirc,qc->oipq", GI, g.astype(np.float64), GTI)
    want = np.einsum("pr,oipq,qc->oirc", G, dU, G)
    assert np.abs(dw - want).max() <= 2e-6 * np.abs(want).max()
    assert np.abs(dw - g).max() <= 1e-5          # G+ G = I up to float32 rounding: nearly the plain STE


def test_fma_emulation_is_single_rounding_on_a_known_case():
    # a*b + c where the product needs more than 24 bits: a separate multiply would round it away
    a, b, c = F32(1 + 2 ** -12), F32(1 + 2 ** -12), F32(-1)
    exact = (1 + 2 ** -12) ** 2 - 1
    assert O._fma32(a, b, c) == F32(exact)
    assert F32(F32(a * b) + c) != F32(exact)
