"""freeze.py of the reference.

Only ``quantize_params`` is on the B200 hot path: the int8 code export with the
arithmetic of MXNet's ``contrib.quantize(out_type="int8")``.  ``quantize_symbol`` / ``FreezeHelper``
rewrite MXNet symbol graphs for the MKLDNN backend through libmxnet's C API; there is no MXNet here,
the reference's README marks that section untested, and SURVEY 8 puts it out of scope -- they raise."""
from collections import OrderedDict

from ... import ops

__all__ = ['quantize_symbol', 'quantize_params', 'FreezeHelper']


def quantize_symbol(sym, excluded_symbols=[], offline_params=[], quantized_dtype='uint8', calib_quantize_op=False):
    raise NotImplementedError("MXQuantizeSymbol is an MXNet symbol-graph rewrite (freeze.py:42-76): out of scope")


def quantize_params(qsym, params):
    """For every name in ``qsym`` (an iterable of argument names, or an object with
    ``list_arguments()``) ending in ``_quantize``: int8 codes + min + max of ``params[name]``
    using ``params[name + "_min"/"_max"]`` (freeze.py:79-109)."""
    import torch
    inputs_name = qsym.list_arguments() if hasattr(qsym, "list_arguments") else list(qsym)
    quantized_params = OrderedDict()
    for name in inputs_name:
        if name.endswith('_quantize'):
            original_name = name[:-len('_quantize')]
            rng = torch.stack([params[original_name + "_min"].reshape(()),
                               params[original_name + "_max"].reshape(())]).float()
            val, out_range = ops.quantize_int8_export(params[original_name], rng)
            quantized_params[name] = val
            quantized_params[name + '_min'] = out_range[0:1]
            quantized_params[name + '_max'] = out_range[1:2]
        elif name in params:
            quantized_params[name] = params[name]
    return quantized_params


class FreezeHelper(object):
    def __init__(self, net, params_filename):
        raise NotImplementedError("FreezeHelper exports MXNet/MKLDNN symbols (freeze.py:134-238): out of scope")
