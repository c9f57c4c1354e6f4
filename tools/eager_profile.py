#!/usr/bin/env python
"""Where the host time of an EAGER forward goes (config 1: cifar_resnet20_v1, online uint8 inputs, batch 128).

The GPU work of this network is ~0.8 ms per forward; the eager step is bound by what Python does per fake-quant
call.  Prints (a) wall time per forward with quantisation enabled / disabled, (b) a cProfile of 200 forwards sorted
by own time, (c) micro-timings of the pieces of one ops.forward_online call."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_configs as BC  # noqa: E402
from quantization.mxnet_b200 import _ffi, ops  # noqa: E402


def wall(fn, n=300):
    for _ in range(30):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return host / n * 1e6, (time.perf_counter() - t0) / n * 1e6


def main():
    cfg_id = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = BC.CONFIGS[cfg_id]
    net = BC.build(cfg, dev)
    net.eval()
    x = torch.randn(*cfg["shape"], device=dev)

    def fwd():
        with torch.no_grad():
            return net(x)
    print("config", cfg_id, "blocks", len(net.collect_quantized_blocks()))
    print("quantised   : host %.1f us, host+gpu %.1f us per forward" % wall(fwd))
    net.disable_quantize()
    print("disabled    : host %.1f us, host+gpu %.1f us per forward" % wall(fwd))
    net.enable_quantize()
    pr = cProfile.Profile()
    fwd()
    pr.enable()
    for _ in range(200):
        fwd()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr, stream=sys.stdout)
    st.sort_stats("tottime").print_stats(28)

    # pieces of one call
    m = net.collect_quantized_blocks()[3]
    a = torch.randn(128, 16, 32, 32, device=dev)

    def t(label, fn, n=20000):
        for _ in range(100):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        dt = (time.perf_counter() - t0) / n * 1e6
        torch.cuda.synchronize()
        print("  %-46s %.2f us" % (label, dt))
    t("dl(large tensor)", lambda: _ffi.dl(a))
    t("dl(small cached tensor)", lambda: _ffi.dl(m.current_input_max))
    t("torch.empty_like", lambda: torch.empty_like(a), 5000)
    t("workspace()", lambda: _ffi.workspace(a.device))
    t("current_stream()", lambda: _ffi.current_stream())
    t("lib.fq_version()", lambda: _ffi.load().fq_version())
    qa = m.quantize_args
    t("ops.forward_online (2 launches, 2 MB)", lambda: ops.forward_online(
        a, qa.in_width, qa.in_signed, ops.LO_ZERO, quantize=True, cur_max=m.current_input_max, qparams=m._fq_qparams), 3000)
    y = torch.empty_like(a)
    t("ops.forward_online (out= given)", lambda: ops.forward_online(
        a, qa.in_width, qa.in_signed, ops.LO_ZERO, quantize=True, cur_max=m.current_input_max, qparams=m._fq_qparams,
        out=y), 3000)
    if hasattr(ops, "InputPlan"):
        plan = ops.InputPlan(a, qa.in_width, qa.in_signed, ops.LO_ZERO, cur_max=m.current_input_max, qparams=m._fq_qparams)
        t("InputPlan.run (same call through the C-side plan)", lambda: plan.run(a), 3000)
        t("InputPlan.run (out= given)", lambda: plan.run(a, y), 3000)
    t("conv2d 16->16 3x3 (framework launch)", lambda: torch.nn.functional.conv2d(a, m.weight[:16, :16] if m.weight.shape[1] >= 16 else m.weight, None, 1, 1), 3000)


if __name__ == "__main__":
    main()
