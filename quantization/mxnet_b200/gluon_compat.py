"""The slice of the Gluon block/parameter vocabulary the reference's Python API leans on, expressed
over torch.nn modules: block names with per-scope type counters (so that
``name.replace("conv", "batchnorm")`` finds the sibling BatchNorm, initialize.py:51-54),
``collect_params`` with ``<block>_<param>`` keys and the nn.* aliases the converter tables use."""
import re
from collections import OrderedDict

import torch
from torch import nn as _tnn


class nn:  # noqa: N801  (mirrors `from mxnet.gluon import nn`)
    Conv2D = _tnn.Conv2d
    Dense = _tnn.Linear
    BatchNorm = _tnn.BatchNorm2d
    Activation = _tnn.ReLU
    HybridSequential = _tnn.Sequential
    Block = _tnn.Module


_HINTS = ((_tnn.Conv2d, "conv"), (_tnn.Linear, "dense"), (_tnn.BatchNorm2d, "batchnorm"), (_tnn.ReLU, "relu"),
          (_tnn.ReLU6, "relu6"), (_tnn.AvgPool2d, "pool"), (_tnn.MaxPool2d, "pool"), (_tnn.AdaptiveAvgPool2d, "pool"),
          (_tnn.Flatten, "flatten"))


class NameScope:
    """Gluon's _BlockScope counters: the i-th block of a type created in a scope is <prefix><hint><i>."""

    def __init__(self, prefix):
        self.prefix = prefix
        self._count = {}

    def name(self, hint):
        i = self._count.get(hint, 0)
        self._count[hint] = i + 1
        return "%s%s%d" % (self.prefix, hint, i)

    def child(self, prefix):
        return NameScope(self.prefix + prefix)

    def __call__(self, module, hint=None):
        if hint is None:
            hint = next((h for t, h in _HINTS if isinstance(module, t)), type(module).__name__.lower())
        module.name = self.name(hint)
        return module


def assign_names(net, prefix="net0_"):
    """Name every leaf block in definition order inside one scope (enough for conv<->batchnorm pairing
    when each conv is followed by its BatchNorm).  Blocks that already carry a name are kept."""
    scope = NameScope(prefix)
    for m in net.modules():
        if len(list(m.children())) == 0 and not hasattr(m, "name"):
            scope(m)
    return net


# gluon parameter suffix -> torch attribute, per block type
_PARAM_ATTRS = {
    _tnn.BatchNorm2d: (("gamma", "weight"), ("beta", "bias"), ("running_mean", "running_mean"),
                       ("running_var", "running_var")),
}
_DEFAULT_ATTRS = (("weight", "weight"), ("bias", "bias"), ("input_max", "input_max"), ("act_max", "act_max"),
                  ("gamma", "gamma"), ("beta", "beta"), ("running_mean", "running_mean"),
                  ("running_var", "running_var"))


def collect_params(net, select=None):
    """OrderedDict ``<block name>_<param>`` -> tensor, optionally filtered by a regex (gluon semantics:
    ``re.match`` on the full name)."""
    out = OrderedDict()
    pat = re.compile(select) if select else None
    for m in net.modules():
        name = getattr(m, "name", None)
        if name is None:
            continue
        attrs = next((a for t, a in _PARAM_ATTRS.items() if isinstance(m, t)), _DEFAULT_ATTRS)
        for suffix, attr in attrs:
            p = getattr(m, attr, None)
            if isinstance(p, torch.Tensor):
                key = "%s_%s" % (name, suffix)
                if pat is None or pat.match(key):
                    out[key] = p
    return out
