"""Activation converters of the reference (convert_act.py): ReLU -> ReLU6, and an optional quantiser
on the activation OUTPUT (off by default there, convert.py:38).  The quantiser runs on the
same kernels as the input path: the range is the mean over samples of the per-sample maximum of the
(non-negative) activation, the divisor carries no 1e-10 and there is no STE."""
import types
from collections import namedtuple

import torch
from torch import nn

from ... import ops

__all__ = ["convert_relu_to_relu6", 'gen_act_converter']

QuantizedArgs = namedtuple("ActQuantizedArgs", "width quantize_act")


def convert_relu_to_relu6(m):
    assert isinstance(m, nn.ReLU)
    plain_relu = m.forward
    m.forward = types.MethodType(lambda self, x: torch.clamp(plain_relu(x), 0., 6.), m)


def _quantised_activation(self, x):
    act = self.origin_forward(x)
    if not (self.enable_quantize and self.quantize_args.quantize_act):
        return act
    if self.current_act_max.device != act.device:
        # a ReLU has no weight to take a device from: its state follows the first activation it sees
        # (convert_model already moves it to the net's device; this covers a block converted on its own)
        self.to(act.device)
    # act >= 0 after a ReLU, so max == absmax (convert_act.py:50)
    ops.input_range(act.detach(), cur_max=self.current_act_max)
    if self.quantize_act:
        max_ = self.act_max.data if self.quantize_act_offline else self.current_act_max
        qp = ops.scale_from_max(max_, self.quantize_args.width, False, ops.LO_ZERO, qparams=self._fq_qparams)
        qp[0:1].copy_(qp[1:2])            # (act.clip(0, max_) / scale).round() * scale: divisor == scale (:53-54)
        act = ops.forward_scalar(act, qp)
    return act


def gen_act_converter(width=8, quantize_act=True):
    def _converter(m):
        assert isinstance(m, nn.ReLU)
        m.quantize_act_offline = False
        m.register_buffer("current_act_max", torch.zeros(1), persistent=False)
        m.register_parameter("act_max", nn.Parameter(torch.zeros(1), requires_grad=False))
        m.register_buffer("_fq_qparams", torch.zeros(4), persistent=False)
        m.origin_forward = m.forward
        m.forward = types.MethodType(_quantised_activation, m)
        m.quantize_args = QuantizedArgs(width=width, quantize_act=quantize_act)
        m.enable_quantize = True
        m.quantize_act = quantize_act
    return _converter
