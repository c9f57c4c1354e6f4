#!/usr/bin/env python
"""Per-kernel times of the tensor-core QConv2D route on ResNet-like layers (batch 32):
pack_input (HBM-bound: 4 B read + 1 B written per element), pack_weight, igemm (tensor-bound), whole layer,
beside the float-code route and a plain cuDNN fp32 convolution.  CUDA events around a graph replay of 20 calls."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantization.mxnet_b200 import ops  # noqa: E402
from quantization.mxnet_b200.nn import Conv2D  # noqa: E402


def timeit(fn, reps=20):
    """GPU time per call in ms: the calls are captured into a CUDA graph and replayed, so that host time per call
    (tens of microseconds through Python) does not hide kernels that are shorter than that."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    full = "--quick" not in sys.argv
    layers = [(32, 256, 56, 256, 3, 1), (32, 64, 56, 64, 3, 1), (32, 128, 28, 128, 3, 1), (32, 512, 14, 512, 3, 1),
              (32, 256, 56, 64, 1, 1), (32, 512, 28, 1024, 1, 2)]
    for n, c, hw, co, k, s in layers:
        conv = Conv2D(co, k, s, k // 2, in_channels=c, quantized=True, input_dtype="int8", weight_dtype="int8").cuda()
        x = torch.randn(n, c, hw, hw, device="cuda")
        with torch.no_grad():
            in_rng, unsigned, w_rng = conv._tensor_core_ranges(x)
            t_pi = timeit(lambda: ops.qconv_pack_input(x, in_rng, k // 2, k // 2))
            t_pw = timeit(lambda: conv._weight_codes(w_rng))
            xq, s_in = ops.qconv_pack_input(x, in_rng, k // 2, k // 2)
            wq, s_w = conv._weight_codes(w_rng)
            t_mm = timeit(lambda: ops.qconv_igemm(xq, wq, None, s_in, s_w, (s, s), 1))
            t_all = timeit(lambda: conv(x))
            t_cudnn = timeit(lambda: torch.nn.functional.conv2d(x, conv.weight, None, s, k // 2))
            t_ref = None
            if full:
                conv.use_tensor_cores = False
                t_ref = timeit(lambda: conv(x), 5)
        ho = (hw + 2 * (k // 2) - k) // s + 1
        ops_ = 2.0 * n * ho * ho * co * c * k * k
        print("N%d C%d %dx%d -> %d, %dx%d/%d: pack_input %.1f us (%.0f GB/s of 5 B/elem), pack_weight %.1f us, igemm %.1f us "
              "= %.0f TOP/s, layer %.1f us; cuDNN fp32 %.1f us%s"
              % (n, c, hw, hw, co, k, k, s, t_pi * 1e3, 5.0 * x.numel() / t_pi / 1e6, t_pw * 1e3, t_mm * 1e3,
                 ops_ / t_mm / 1e9, t_all * 1e3, t_cudnn * 1e3,
                 "" if t_ref is None else ", float-code route %.1f us" % (t_ref * 1e3)), flush=True)


if __name__ == "__main__":
    main()
