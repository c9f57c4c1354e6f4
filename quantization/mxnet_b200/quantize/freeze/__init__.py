"""quantize.freeze: BN merging and integer export."""
from .freeze import FreezeHelper, export_quantized, quantize_params, quantize_symbol
from .merge_bn import merge_bn

__all__ = ["merge_bn", "quantize_symbol", "quantize_params", "FreezeHelper", "export_quantized"]
