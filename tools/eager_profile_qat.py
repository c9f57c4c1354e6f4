#!/usr/bin/env python
"""Host-side profile of the EAGER config-3 QAT step (mobilenetv2_1.0, notebook converters, batch 128): wall time per
step quantised / disabled, and a cProfile of 60 steps sorted by own time."""
import cProfile
import os
import pstats
import sys
import time

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_configs as BC  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    cfg = BC.CONFIGS[3]
    net = BC.build(cfg, dev)
    X = torch.randn(*cfg["shape"], device=dev)
    y = torch.randint(0, cfg["classes"], (cfg["shape"][0],), device=dev)
    net.train()
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.eval()
    net.quantize_input(enable=True, online=True)
    with torch.no_grad():
        net(X)
    net.update_ema()
    net.quantize_input(enable=True, online=False)
    params = [p for p in net.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-6, fused=True)
    loss_fn = nn.CrossEntropyLoss()

    def step():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(net(X), y)
        loss.backward()
        net.update_ema()
        opt.step()
        return loss

    def wall(n=40):
        for _ in range(8):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            step()
        host = time.perf_counter() - t0
        torch.cuda.synchronize()
        return host / n * 1e3, (time.perf_counter() - t0) / n * 1e3
    print("quantised: host %.2f ms, host+gpu %.2f ms per step" % wall())
    net.disable_quantize()
    print("disabled : host %.2f ms, host+gpu %.2f ms per step" % wall())
    net.enable_quantize()
    step()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(60):
        step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr, stream=sys.stdout).sort_stats("tottime").print_stats(32)


if __name__ == "__main__":
    main()
