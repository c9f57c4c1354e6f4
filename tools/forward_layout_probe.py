#!/usr/bin/env python
"""Framework forward of the e2e leg (torch mobilenet1.0, fp32, TF32 off, batch 128 at 224x224): NCHW vs channels_last,
cudnn.benchmark on.  The histogram / minmax kernels are order-independent, so a dense channels_last activation could
be histogrammed as it lies if the framework forward were faster that way."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    net = bench.build_net(dev)
    net.disable_quantize()
    net.eval()
    x = torch.randn(128, 3, 224, 224, device=dev)
    for label, fmt in (("NCHW", torch.contiguous_format), ("channels_last", torch.channels_last)):
        n2 = net.to(memory_format=fmt)
        xx = x.contiguous(memory_format=fmt)
        with torch.no_grad():
            for _ in range(5):
                n2(xx)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                n2(xx)
            b.record()
            torch.cuda.synchronize()
        print("%-14s forward %.3f ms per batch of 128" % (label, a.elapsed_time(b) / 10), flush=True)


if __name__ == "__main__":
    main()
