"""B200-native fake-quantization hot path with the Python API of hey-yahei/Quantization.MXNet."""
__version__ = "0.1.0"
