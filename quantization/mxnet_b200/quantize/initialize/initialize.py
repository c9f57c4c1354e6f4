"""initialize.py of the reference: zero the input ranges; for fake-BN convs adopt the sibling
BatchNorm's gamma/beta/mean/var (found through the gluon naming convention) and grow a bias."""
import torch
from torch import nn

from ...gluon_compat import collect_params

__all__ = ["qparams_init"]


def qparams_init(net, conv_name="conv", bn_name="batchnorm"):
    blocks = net.collect_quantized_blocks()
    params = collect_params(net)

    for m in blocks:
        # If fake bn, initialize the related params from the sibling batchnorm (initialize.py:46-70)
        if isinstance(m, nn.Conv2d) and hasattr(m, "gamma"):
            name = m.name
            bn = name.replace(conv_name, bn_name)
            with torch.no_grad():
                m.gamma.copy_(params[bn + "_gamma"])
                m.beta.copy_(params[bn + "_beta"])
                m.running_mean.copy_(params[bn + "_running_mean"])
                m.running_var.copy_(params[bn + "_running_var"])
            # Enable bias if need
            if m.bias is None:
                m.bias = nn.Parameter(torch.zeros(m.out_channels, dtype=m.weight.dtype, device=m.weight.device))

        if type(m) in (nn.Conv2d, nn.Linear) and m.quantize_args.quantize_input:
            with torch.no_grad():
                m.input_max.zero_()
        if type(m) == nn.ReLU and m.quantize_args.quantize_act:
            with torch.no_grad():
                m.act_max.zero_()
