"""Net-level entry points of the reference's ``quantize/convert/convert.py``.

``convert_model`` visits the blocks children-first, hands each one whose EXACT type has a converter to
that converter, and then equips ``net`` with the six control methods of the reference
(``update_ema``, ``collect_quantized_blocks``, ``quantize_input``, ``enable_quantize``,
``disable_quantize``, ``fix_params``; convert.py:66-121).  Here they live on one controller class and
are bound to the net, and the EMA of every layer is a single kernel launch over packed state.
"""
import types

import torch
from torch import nn

from ... import ops
from . import _state
from .convert_act import convert_relu_to_relu6, gen_act_converter
from .convert_bn import bypass_bn
from .convert_conv2d import gen_conv2d_converter, prequantize_weights, sync_pending_ranges, sync_pending_stats
from .convert_dense import gen_dense_converter

__all__ = ["convert_model", "convert_to_relu6", 'default_convert_fn']

# Activation / BatchNorm are left alone by default, as in the reference (its alternatives are
# convert_relu_to_relu6 / gen_act_converter() and bypass_bn).
default_convert_fn = {nn.Conv2d: gen_conv2d_converter(), nn.Linear: gen_dense_converter(),
                      nn.ReLU: None, nn.BatchNorm2d: None}

_WEIGHTED = (nn.Linear, nn.Conv2d)
_QUANTISABLE = _WEIGHTED + (nn.ReLU,)


class _Controls:
    """Bound onto the converted net by :func:`convert_model`."""

    def collect_quantized_blocks(self):
        found = []
        self.apply(lambda m: found.append(m) if type(m) in _QUANTISABLE and hasattr(m, 'quantize_args') else None)
        return found

    def update_ema(self, momentum=0.9):
        """state <- (1 - momentum) * current + momentum * state for input_max, act_max and the fake-BN
        running statistics (convert.py:66-78)."""
        blocks = self.collect_quantized_blocks()
        bucket = getattr(self, "_fq_grad_bucket", None)
        if bucket is not None and bucket.takes_over_ema():
            # data-parallel training step: the ranges travel with the gradient all-reduce and the EMA is applied
            # right after it (nothing reads input_max before the next forward)
            bucket.defer_ema(momentum)
            return
        sync_pending_ranges(blocks)       # data parallel only: shard-local ranges -> global-batch ranges
        arenas = _state.pack(self, blocks)
        sync_pending_stats(blocks, arenas)     # data parallel only: shard-local fake-BN statistics -> global batch
        # one launch per kind for the whole net (the reference: one or two ops per block, convert.py:68-78)
        for state_name, _, scalar_cur in _state.KINDS:
            arena = arenas.get(state_name)
            if arena is not None:
                ops.ema_update(arena["state"], arena["current"], momentum, scalar_cur=scalar_cur)

    def quantize_input(self, enable=True, online=True):
        """Switch input (or activation) quantisation on/off and between online and offline ranges."""
        for m in self.collect_quantized_blocks():
            if type(m) in _WEIGHTED:
                assert (not enable) or m.quantize_args.quantize_input
                m.quantize_input, m.quantize_input_offline = enable, not online
            elif type(m) == nn.ReLU:
                assert (not enable) or m.quantize_args.quantize_act
                m.quantize_act, m.quantize_act_offline = enable, not online

    def enable_quantize(self):
        for m in self.collect_quantized_blocks():
            m.enable_quantize = True

    def disable_quantize(self):
        for m in self.collect_quantized_blocks():
            m.enable_quantize = False

    def fix_params(self):
        """Ask every converted Conv2D to cache its quantised weight (and folded bias) at its next forward.
        Dense blocks are not touched -- the reference does not fix them either (convert.py:117-121)."""
        for m in self.collect_quantized_blocks():
            if isinstance(m, nn.Conv2d):
                m.fixed_params = 0


_CONTROL_NAMES = ("update_ema", "collect_quantized_blocks", "quantize_input", "enable_quantize", "disable_quantize",
                  "fix_params")


def _before_net_forward(net, args):
    # The reference's convolutions are plain fp32; torch would run them in TF32 on this GPU unless told otherwise
    # (cudnn.allow_tf32 defaults to True), which breaks the 1e-5 logits contract and the fake-BN statistics.
    net.__dict__["_fq_tf32_saved"] = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    if getattr(net, "force_fp32_framework_ops", True):
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    if getattr(net, "batch_weight_paths", True):           # set to False to force the per-block launches
        prequantize_weights(net, net._fq_hook_blocks)


def _after_net_forward(net, args, output):
    saved = net.__dict__.pop("_fq_tf32_saved", None)
    if saved is not None:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    for m in net._fq_hook_blocks:                 # a block the forward did not reach must not keep a stale result
        m.__dict__.pop("_fq_pre", None)


def convert_model(net, exclude=[], convert_fn=default_convert_fn, custom_fn={}):
    """Convert ``net`` in place to its simulated-quantisation version.

    net        : torch.nn.Module
    exclude    : blocks to leave untouched
    convert_fn : {block type: converter(block) -> None}; looked up with the block's exact type
    custom_fn  : {block instance: converter}; wins over ``convert_fn``
    Returns None, like the reference (whose docstring promises the net).
    """
    skip = set(map(id, exclude))

    def visit(m):
        if id(m) in skip:
            return
        fn = custom_fn[m] if m in custom_fn else convert_fn.get(type(m))
        if fn is not None:
            fn(m)
    net.apply(visit)
    # converters of weightless blocks (ReLU) cannot know the net's device: put their state where the net lives
    first = next(net.parameters(), None)
    if first is not None:
        for m in net.modules():
            if type(m) == nn.ReLU and hasattr(m, "quantize_args") and m.act_max.device != first.device:
                m.to(first.device)
    for name in _CONTROL_NAMES:
        setattr(net, name, types.MethodType(getattr(_Controls, name), net))
    # B200 execution detail (not part of the reference's API): the weight paths of ALL blocks run as one
    # multi-tensor launch at the start of every net-level forward instead of ~50 tiny launches spread over it.
    # Weights do not depend on activations, so the results are the same; a block called on its own still
    # takes its per-block path.
    net._fq_hook_blocks = net.collect_quantized_blocks()   # blocks converted later simply keep their per-block path
    if not getattr(net, "_fq_weight_hooks", False):
        net.register_forward_pre_hook(_before_net_forward)
        net.register_forward_hook(_after_net_forward, always_call=True)
        net._fq_weight_hooks = True
    # every range / statistic of the net in a few arenas, built now so that pointers are stable before any
    # CUDA-graph capture, and rebuilt whenever the net is moved
    _state.install(net)


def convert_to_relu6(net, exclude=[]):
    """Turn every ReLU of ``net`` (except those in ``exclude``) into a ReLU6; returns the net."""
    skip = set(map(id, exclude))
    return net.apply(lambda m: convert_relu_to_relu6(m) if isinstance(m, nn.ReLU) and id(m) not in skip else None)
