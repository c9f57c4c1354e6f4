"""``merge_bn`` of the reference (freeze/merge_bn.py): fold every BatchNorm into the convolution in front of
it for good -- W' = (W * gamma) / sqrt(var + 1e-10), b' = gamma * (b - mean) / sqrt(var + 1e-10) + beta, the
same kernel as the fake-BN fold of the converted forward -- and turn the BatchNorm into an identity."""
import torch
from torch import nn

from ... import ops
from ...gluon_compat import collect_params
from ..convert.convert_bn import bypass_bn

__all__ = ['merge_bn']


def merge_bn(net, conv_name="conv", bn_name="batchnorm", exclude=[]):
    """conv_name / bn_name: the keywords that turn a convolution's gluon-style name into its BatchNorm's
    (``..._conv3`` -> ``..._batchnorm3``); exclude: convolutions (and BatchNorms) to leave alone."""
    skip = set(map(id, exclude))
    params = collect_params(net)
    for conv in [m for m in net.modules() if isinstance(m, nn.Conv2d)]:
        # the reference refuses to merge into a convolution that carries a fake BN (merge_bn.py:48)
        assert not hasattr(conv, "gamma"), "Don't merge bn to a conv with fake bn! ({})".format(conv.name)
        bn = conv.name.replace(conv_name, bn_name)
        if id(conv) in skip or bn + "_gamma" not in params:
            continue
        print("Merge {} to {}".format(bn, conv.name))
        with torch.no_grad():
            w, b, _ = ops.quant_weight(conv.weight.data, 1, 0, params[bn + "_gamma"].data, params[bn + "_beta"].data,
                                       params[bn + "_running_mean"], params[bn + "_running_var"],
                                       None if conv.bias is None else conv.bias.data)
            conv.weight.copy_(w)
            if conv.bias is None:
                conv.bias = nn.Parameter(b)
            else:
                conv.bias.copy_(b)
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d) and id(m) not in skip:
            bypass_bn(m)
