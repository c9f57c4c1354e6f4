"""The reference's ``quantize`` package over the B200 kernels: ``convert``, ``initialize``, ``freeze``,
``distribution_calibrate`` and the qparam helpers of ``utils``."""
from . import convert, distribution_calibrate, freeze, initialize
from .utils import collect_qparams, print_all_qparams

__all__ = ["convert", "initialize", "freeze", "distribution_calibrate", "collect_qparams", "print_all_qparams"]
