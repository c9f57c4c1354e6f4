"""``bypass_bn`` of the reference (convert_bn.py:32-36): a BatchNorm whose statistics were folded into
the preceding convolution (fake-BN / merge-BN) becomes the identity."""
import types

from torch import nn

__all__ = ['bypass_bn']


def _identity(self, x, *unused, **unused_kw):
    return x


def bypass_bn(m):
    assert isinstance(m, nn.BatchNorm2d)
    m.forward = types.MethodType(_identity, m)
