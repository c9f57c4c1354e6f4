"""freeze.py of the reference.

Only ``quantize_params`` is on the B200 hot path: the int8 code export with the
arithmetic of MXNet's ``contrib.quantize(out_type="int8")``.  ``quantize_symbol`` / ``FreezeHelper``
rewrite MXNet symbol graphs for the MKLDNN backend through libmxnet's C API; there is no MXNet here,
the reference's README marks that section untested, and SURVEY 8 puts it out of scope -- they raise."""
from collections import OrderedDict

from ... import ops

__all__ = ['quantize_symbol', 'quantize_params', 'FreezeHelper', 'export_quantized']


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


def export_quantized(net, dtype="int8"):
    """Framework-neutral replacement for the MXNet-symbol export: for every converted
    Conv2D / Dense block an entry with the zero-centred int8 weight codes and their range (the arithmetic
    of ``contrib.quantize``, freeze.py:100-103), the folded bias, and the calibrated input threshold that
    FreezeHelper._set_min_max would write as ``max_calib_range`` (freeze.py:191-208).

    Returns an OrderedDict ``block name -> dict`` of device tensors / Python scalars; nothing is read back
    to the host except the scalar thresholds."""
    import torch
    from torch import nn
    assert dtype == "int8"
    out = OrderedDict()
    for m in net.collect_quantized_blocks():
        if not isinstance(m, (nn.Conv2d, nn.Linear)):
            continue
        w, b = m.weight.detach(), (None if m.bias is None else m.bias.detach())
        if getattr(m, "fixed_params", 1) != 1 and getattr(m.quantize_args, "fake_bn", False):
            w, b, _ = ops.quant_weight(w, 1, 0, m.gamma.data, m.beta.data, m.running_mean.data, m.running_var.data, b)
        mx = ops.absmax_rows(w, 1)
        codes, rng = ops.quantize_int8_export(w, torch.cat([-mx, mx]))
        entry = {"weight_quantize": codes, "weight_min": rng[0:1], "weight_max": rng[1:2], "bias": b}
        if m.quantize_args.quantize_input:
            th = float(m.input_max.detach().reshape(-1)[0])
            entry["min_calib_range"] = -th if m.quantize_args.in_signed else 0.0
            entry["max_calib_range"] = th
        out[getattr(m, "name", str(id(m)))] = entry
    return out
