"""Dense converter (convert_dense.py of the reference) over torch.nn.Linear.

Reference quirks kept: signed inputs are still clipped at 0 because no clip_min is passed
(convert_dense.py:49); "group" maps to "channel" (:83-84); weights are re-quantised on every forward
(there is no fixed_params for Dense)."""
import types
from collections import namedtuple

import torch
from torch import nn

from ... import ops
from .convert_conv2d import _input_path, _weight_path, _range_only

__all__ = ['gen_dense_converter']

QuantizedArgs = namedtuple("DenseQuantizedArgs", "in_signed in_width wt_width quantize_input quant_type")


def _dense_forward(self, x):
    qa = self.quantize_args
    weight, bias = self.weight, self.bias
    if self.enable_quantize:
        if qa.quantize_input:
            if x.dim() != 2:
                raise NotImplementedError("quantised Dense expects a flattened (N, in_units) input")
            if self.quantize_input:
                x = _input_path(x, self, ops.LO_ZERO)
            else:
                _range_only(x.detach(), self)
        pre = self.__dict__.pop("_fq_pre", None)       # set by the net-level multi-tensor launch, used once
        if pre is not None:
            weight_q = pre[0]
        else:
            rows = self.out_features if qa.quant_type == 'channel' else 1
            weight_q, _ = _weight_path(weight, None, None, None, None, None, rows, qa.wt_width)
    else:
        self.__dict__.pop("_fq_pre", None)
        weight_q = weight
    return self.origin_forward(x, weight_q, bias)


def _add_quantize_input_params(m):
    m.quantize_input_offline = False
    dev = m.weight.device
    # non-persistent buffers follow .cuda()/.to() but stay out of the state dict, like the reference's
    # plain attribute (current_input_max is not checkpointed)
    m.register_buffer("current_input_max", torch.zeros(1, dtype=torch.float32, device=dev), persistent=False)
    m.register_parameter("input_max", nn.Parameter(torch.zeros(1, dtype=torch.float32, device=dev),
                                                   requires_grad=False))
    m.register_buffer("_fq_qparams", torch.zeros(4, dtype=torch.float32, device=dev), persistent=False)


def gen_dense_converter(weight_width=8, input_signed=False, input_width=8, quantize_input=True, quant_type='layer'):
    if quant_type == "group":
        quant_type = "channel"

    def _converter(m):
        assert isinstance(m, nn.Linear)

        if quantize_input:
            _add_quantize_input_params(m)
        m.origin_forward = types.MethodType(lambda self, x, w, b: nn.functional.linear(x, w, b), m)
        m.forward = types.MethodType(_dense_forward, m)
        m.quantize_args = QuantizedArgs(in_signed=input_signed, in_width=input_width, wt_width=weight_width,
                                        quantize_input=quantize_input, quant_type=quant_type)
        m.enable_quantize = True
        m.quantize_input = quantize_input
    return _converter
