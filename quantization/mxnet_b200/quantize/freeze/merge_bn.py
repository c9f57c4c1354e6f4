"""merge_bn.py of the reference: fold every BatchNorm into its convolution for good and turn the
BatchNorm into an identity.  The fold arithmetic (eps = 1e-10, multiply then divide) runs in the same
CUDA kernel as the fake-BN fold of the converted forward."""
import types

import torch
from torch import nn

from ... import ops
from ...gluon_compat import collect_params

__all__ = ['merge_bn']


def _bypass_bn(net, exclude=[]):
    exclude_ids = set(id(b) for b in exclude)

    def _forward(self, x, *args, **kwargs):
        return x

    def _bypass(m):
        if isinstance(m, nn.BatchNorm2d) and id(m) not in exclude_ids:
            m.forward = types.MethodType(_forward, m)
    net.apply(_bypass)


def _merge_bn(net, conv_name="conv", bn_name="batchnorm", exclude=[]):
    exclude_ids = set(id(b) for b in exclude)
    conv_lst = []

    def _collect_conv(m):
        if isinstance(m, nn.Conv2d):
            assert not hasattr(m, "gamma"), "Don't merge bn to a conv with fake bn! ({})".format(m.name)
            conv_lst.append(m)
    net.apply(_collect_conv)

    bn_names = [c.name.replace(conv_name, bn_name) for c in conv_lst]
    for conv, bn in zip(conv_lst, bn_names):
        params = collect_params(net, bn + "_")
        if len(params.keys()) != 0 and id(conv) not in exclude_ids:
            print("Merge {} to {}".format(bn, conv.name))
            gamma = params[bn + "_gamma"]
            beta = params[bn + "_beta"]
            mean = params[bn + "_running_mean"]
            var = params[bn + "_running_var"]
            with torch.no_grad():
                w, b, _ = ops.quant_weight(conv.weight.data, 1, 0, gamma.data, beta.data, mean, var,
                                           None if conv.bias is None else conv.bias.data)
                conv.weight.copy_(w)
                if conv.bias is None:
                    conv.bias = nn.Parameter(b)
                else:
                    conv.bias.copy_(b)


def merge_bn(net, conv_name="conv", bn_name="batchnorm", exclude=[]):
    """Merge all batchnorm to convolution (names follow the gluon convention, see gluon_compat)."""
    _merge_bn(net, conv_name, bn_name, exclude)
    _bypass_bn(net, exclude)
