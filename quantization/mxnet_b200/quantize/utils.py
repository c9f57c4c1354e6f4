"""Qparam helpers of the reference's ``quantize/utils.py``."""
from collections import OrderedDict

from ..gluon_compat import collect_params

__all__ = ['collect_qparams', 'print_all_qparams']


def collect_qparams(net):
    """OrderedDict of every quantisation range of ``net`` -- the parameters whose gluon-style name ends in
    ``_min`` or ``_max`` (``<block>_input_max``, ``<block>_act_max``) -- keyed by that name."""
    return OrderedDict((name, p) for name, p in collect_params(net).items() if name.endswith(("_min", "_max")))


def print_all_qparams(net):
    for name, p in collect_qparams(net).items():
        print("{}:\t\t{:+.4f}".format(name, float(p.reshape(-1)[0])))
