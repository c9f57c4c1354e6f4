"""quantize.initialize"""
from .initialize import qparams_init

__all__ = ["qparams_init"]
