"""convert_act.py of the reference: ReLU -> ReLU6 and the optional activation-output quantiser.

The activation quantiser is disabled by default in the reference (convert.py:38) and is listed as a
"next" row; it reuses the same kernels: range = mean of per-sample maxima of a
non-negative activation, no epsilon on the divisor, no STE."""
import types
from collections import namedtuple

import torch
from torch import nn

from ... import ops

__all__ = ["convert_relu_to_relu6", 'gen_act_converter']

QuantizedArgs = namedtuple("ActQuantizedArgs", "width quantize_act")


def _relu6_forward(self, x):
    return torch.clamp(self.origin_relu(x), 0., 6.)


def convert_relu_to_relu6(m):
    assert isinstance(m, nn.ReLU)
    m.origin_relu = m.forward
    m.forward = types.MethodType(_relu6_forward, m)


def _act_forward(self, x):
    act = self.origin_forward(x)
    if self.enable_quantize and self.quantize_args.quantize_act:
        # F.max(act, axis=(1,2,3)).mean(): act >= 0 after ReLU, so max == absmax (convert_act.py:50)
        ops.input_range(act.detach(), cur_max=self.current_act_max)
        if self.quantize_act:
            max_ = self.act_max.data if self.quantize_act_offline else self.current_act_max
            # scale = max_ / (2**w - 1); (act.clip(0, max_) / scale).round() * scale  -- no 1e-10 here (:53-54)
            qp = ops.scale_from_max(max_, self.quantize_args.width, False, ops.LO_ZERO, qparams=self._fq_qparams)
            qp[0:1].copy_(qp[1:2])
            act = ops.forward_scalar(act, qp)
    return act


def _add_quantize_act_params(m):
    m.quantize_act_offline = False
    m.register_buffer("current_act_max", torch.zeros(1, dtype=torch.float32), persistent=False)
    m.register_parameter("act_max", nn.Parameter(torch.zeros(1, dtype=torch.float32), requires_grad=False))
    m.register_buffer("_fq_qparams", torch.zeros(4, dtype=torch.float32), persistent=False)


def gen_act_converter(width=8, quantize_act=True):
    def _converter(m):
        assert isinstance(m, nn.ReLU)

        _add_quantize_act_params(m)

        m.origin_forward = m.forward
        m.forward = types.MethodType(_act_forward, m)
        m.quantize_args = QuantizedArgs(width=width, quantize_act=quantize_act)
        m.enable_quantize = True
        m.quantize_act = quantize_act
    return _converter
