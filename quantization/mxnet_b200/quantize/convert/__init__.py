"""quantize.convert: converter factories, the net-level driver and the fake-quant primitive."""
from . import convert_act, convert_bn, convert_conv2d, convert_dense, ste_func, wino_matrix
from .convert import convert_model, convert_to_relu6, default_convert_fn
from .convert_act import convert_relu_to_relu6, gen_act_converter
from .convert_bn import bypass_bn
from .convert_conv2d import gen_conv2d_converter
from .convert_dense import gen_dense_converter
from .ste_func import LinearQuantizeSTE

__all__ = ["gen_conv2d_converter", "gen_dense_converter", "gen_act_converter", "convert_relu_to_relu6", "bypass_bn",
           "convert_model", "convert_to_relu6", "default_convert_fn", "LinearQuantizeSTE"]
