"""convert.py of the reference: walk the net, patch blocks, bolt the control methods onto ``net``."""
import types

import torch
from torch import nn

from ... import ops
from .convert_conv2d import gen_conv2d_converter, sync_pending_ranges
from .convert_dense import gen_dense_converter
from .convert_act import gen_act_converter, convert_relu_to_relu6
from .convert_bn import bypass_bn

__all__ = ["convert_model", "convert_to_relu6", 'default_convert_fn']

default_convert_fn = {
    nn.Conv2d: gen_conv2d_converter(),
    nn.Linear: gen_dense_converter(),
    nn.ReLU: None,  # convert_relu_to_relu6,  # gen_act_converter(),
    nn.BatchNorm2d: None  # bypass_bn
}

_QBLOCK_TYPES = (nn.Linear, nn.Conv2d, nn.ReLU)


def _pack_states(blocks, attr, cur_attr):
    """Keep every block's (1,) state and its (1,) current value in two contiguous device vectors so
    that the EMA of ALL layers is one launch.  Re-packed lazily after .cuda()/.to()."""
    ms = [m for m in blocks if getattr(m, attr, None) is not None]
    if not ms:
        return None, None, ms
    dev = getattr(ms[0], attr).device
    base = getattr(ms[0], attr).data
    packed = all(getattr(m, attr).data.data_ptr() == base.data_ptr() + 4 * i and
                 getattr(m, cur_attr).data_ptr() == getattr(ms[0], cur_attr).data_ptr() + 4 * i
                 for i, m in enumerate(ms))
    owner = getattr(ms[0], "_fq_arena", None)
    if not (packed and owner is not None and owner[0].device == dev and owner[0].numel() == len(ms)):
        state = torch.cat([getattr(m, attr).data.reshape(1).to(dev) for m in ms])
        cur = torch.cat([getattr(m, cur_attr).reshape(1).to(dev) for m in ms])
        for i, m in enumerate(ms):
            getattr(m, attr).data = state[i:i + 1]
            setattr(m, cur_attr, cur[i:i + 1])
            m._fq_arena = (state, cur)
        owner = (state, cur)
    return owner[0], owner[1], ms


def convert_model(net, exclude=[], convert_fn=default_convert_fn, custom_fn={}):
    """
    Convert the model to the one with simulated quantization.
    :param net: torch.nn.Module
        The net to convert.
    :param exclude: list of torch.nn.Module
        Blocks that want to exclude.
    :param convert_fn: dict with (module type, func) key-value pairs
        `func`: function `func(module) -> None`, applied to blocks whose EXACT type is the key.
    :param custom_fn: dict with (module instance, func) pairs overriding `convert_fn`.
    """
    exclude_ids = set(id(b) for b in exclude)

    # Convert network
    def _convert(m):
        if id(m) not in exclude_ids:
            fn = custom_fn[m] if m in custom_fn else convert_fn.get(type(m))
            if fn is not None:
                fn(m)
    net.apply(_convert)

    # Add method to update ema for `input_max` in convs (convert.py:66-78)
    def _update_ema(self, momentum=0.9):
        blocks = self.collect_quantized_blocks()
        sync_pending_ranges(blocks)     # data parallel only: shard-local ranges -> global-batch ranges
        # if quantize input: every layer's scalar EMA in ONE launch
        state, cur, _ = _pack_states(blocks, "input_max", "current_input_max")
        if state is not None:
            ops.ema_update(state, cur, momentum, scalar_cur=True)
        # if quantize activation
        state, cur, _ = _pack_states(blocks, "act_max", "current_act_max")
        if state is not None:
            ops.ema_update(state, cur, momentum, scalar_cur=True)
        # if fake bn
        for qblocks in blocks:
            if getattr(qblocks, "running_mean", None) is not None and getattr(qblocks, "current_mean", None) is not None:
                ops.ema_update(qblocks.running_mean.data, qblocks.current_mean, momentum, scalar_cur=False)
            if getattr(qblocks, "running_var", None) is not None and getattr(qblocks, "current_var", None) is not None:
                ops.ema_update(qblocks.running_var.data, qblocks.current_var, momentum, scalar_cur=False)
    net.update_ema = types.MethodType(_update_ema, net)

    # Add a method to collect all quantized convolution blocks
    def _collect_quantized_blocks(self):
        blocks = []

        def _collect_blocks(m):
            if type(m) in _QBLOCK_TYPES and hasattr(m, 'quantize_args'):
                blocks.append(m)
        net.apply(_collect_blocks)
        return blocks
    net.collect_quantized_blocks = types.MethodType(_collect_quantized_blocks, net)

    # Add method to control the mode of input quantization as online or offline
    def _quantize_input(self, enable=True, online=True):
        for qblocks in self.collect_quantized_blocks():
            if type(qblocks) in (nn.Linear, nn.Conv2d):
                assert (not enable) or qblocks.quantize_args.quantize_input
                qblocks.quantize_input = enable
                qblocks.quantize_input_offline = not online
            elif type(qblocks) == nn.ReLU:
                assert (not enable) or qblocks.quantize_args.quantize_act
                qblocks.quantize_act = enable
                qblocks.quantize_act_offline = not online
    net.quantize_input = types.MethodType(_quantize_input, net)

    # Add method to control enable/disable quantization
    def _enable_quantize(self):
        for qblocks in self.collect_quantized_blocks():
            qblocks.enable_quantize = True

    def _disable_quantize(self):
        for qblocks in self.collect_quantized_blocks():
            qblocks.enable_quantize = False
    net.enable_quantize = types.MethodType(_enable_quantize, net)
    net.disable_quantize = types.MethodType(_disable_quantize, net)

    # Add method to fixed parameters(weights and bias) -- Conv2D only, as in the reference
    def _fix_params(self):
        for m in net.collect_quantized_blocks():
            if isinstance(m, nn.Conv2d):
                m.fixed_params = 0
    net.fix_params = types.MethodType(_fix_params, net)


def convert_to_relu6(net, exclude=[]):
    """Convert ReLUs in net to ReLU6."""
    exclude_ids = set(id(b) for b in exclude)

    def _convert_to_relu6(m):
        if isinstance(m, nn.ReLU) and id(m) not in exclude_ids:
            convert_relu_to_relu6(m)
    return net.apply(_convert_to_relu6)
