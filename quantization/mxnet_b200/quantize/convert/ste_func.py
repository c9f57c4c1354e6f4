"""The fake-quant primitive (ste_func.py:30-44 of the reference) as a torch.autograd.Function whose
forward is one hand-written CUDA kernel and whose backward is the reference's identity."""
import numpy as np
import torch

from ... import ops

__all__ = ['LinearQuantizeSTE']


class _ScalarSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, d, s, lo, hi, clip):
        return ops.forward_scalar_host(x, d, s, lo, hi, clip)

    @staticmethod
    def backward(ctx, dy):                      # ste_func.py:43-44
        return dy, None, None, None, None, None


class _DeviceQParamSTE(torch.autograd.Function):
    """Scalar quantiser whose {d, s, lo, hi} live on the device (no .asscalar() round trip)."""

    @staticmethod
    def forward(ctx, x, qparams):
        return ops.forward_scalar(x, qparams)

    @staticmethod
    def backward(ctx, dy):
        return dy, None


class _RowsSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        return ops.forward_rows(x, scale)

    @staticmethod
    def backward(ctx, dy):
        return dy, None


def _host_divisor(scale):
    """(d, s) as the MXNet scalar ops finally see them: DType(scale + 1e-10), DType(scale)."""
    if isinstance(scale, np.float32) and ops.get_promotion() == "nep50":
        return float(np.float32(scale + np.float32(1e-10))), float(scale)
    s64 = float(scale)                           # python float / numpy.float64 / legacy float32 -> float64
    return float(np.float32(s64 + 1e-10)), float(np.float32(s64))


class LinearQuantizeSTE(object):
    """``LinearQuantizeSTE(scale, clip_max=None, clip_min=None)(x)``

    forward : ``round(x / (scale + 1e-10)) * scale`` or, with ``clip_max``,
              ``round(clip(x, clip_min, clip_max) / (scale + 1e-10)) * scale``; ``clip_min`` defaults to 0.
    backward: ``dy`` unchanged.  ``scale`` is a constant: no gradient flows into it.

    ``scale`` may be a host scalar (inputs) or a tensor with one entry per leading-axis row of ``x`` /
    a single entry (weights), exactly the two call shapes the reference uses.
    """

    def __init__(self, scale, clip_max=None, clip_min=None):
        self.clip_max = clip_max
        self.clip_min = clip_min if clip_min is not None else 0.
        self.scale = scale

    def __call__(self, x):
        return self.forward(x)

    def forward(self, x):
        if isinstance(self.scale, torch.Tensor):
            if self.clip_max is not None:
                raise NotImplementedError("tensor scale with clipping is not a call shape of the reference")
            scale = self.scale.detach()
            assert scale.numel() in (1, x.shape[0]), "scale must be per-layer or per leading-axis row"
            return _RowsSTE.apply(x, scale)
        d, s = _host_divisor(self.scale)
        if self.clip_max is None:
            return _ScalarSTE.apply(x, d, s, 0.0, 0.0, False)
        return _ScalarSTE.apply(x, d, s, float(self.clip_min), float(self.clip_max), True)

    def backward(self, dy):
        return dy
