"""convert_bn.py of the reference: BatchNorm -> identity (used with fake-BN / merge-BN)."""
import types

from torch import nn

__all__ = ['bypass_bn']


def bypass_bn(m):
    assert isinstance(m, nn.BatchNorm2d)

    def _forward(self, x, *args, **kwargs):
        return x
    m.forward = types.MethodType(_forward, m)
