"""quantize/utils.py of the reference (collect_qparams :30-46, print_all_qparams :49-52)."""
from collections import OrderedDict

from ..gluon_compat import collect_params

__all__ = ['collect_qparams', 'print_all_qparams']


def collect_qparams(net):
    """All ``*_min`` / ``*_max`` quantisation parameters, keyed by gluon-style name."""
    ret = OrderedDict()
    quant_params = collect_params(net, ".*[min|max]")
    for param in quant_params:
        if param.endswith(("_min", "_max")):
            ret[param] = quant_params[param]
    return ret


def print_all_qparams(net):
    qparams = collect_qparams(net)
    for param in qparams:
        print("{}:\t\t{:+.4f}".format(param, float(qparams[param].reshape(-1)[0])))
