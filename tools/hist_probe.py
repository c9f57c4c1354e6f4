"""Times the multi-tensor histogram launch on the config-2 layer inputs (27 tensors, 2.56 GB) and each layer's
single-tensor launch.  The launch shape in csrc/fq_calib.cu (256 threads x 3 blocks per SM, tiles round-robin)
was chosen with this probe; gpurun_out/hist_probe*.log hold the sweeps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from quantization.mxnet_b200 import ops  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3


def main():
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = False
    net = bench.build_net(dev)
    net.disable_quantize()
    X = torch.randn(bench.BATCH, 3, 224, 224, generator=torch.Generator().manual_seed(7)).to(dev)
    acts = bench.capture_layer_inputs(net, X)
    n = sum(a.numel() for a in acts)
    L, B = len(acts), bench.BINS
    minmax = torch.zeros(L, 2, device=dev)
    for i, a in enumerate(acts):
        ops.minmax(a, out=minmax[i])
    print("elements %d, zeros %.3f" % (n, sum(int((a == 0).sum()) for a in acts) / n))
    counts = torch.zeros(L, B + 1, dtype=torch.int64, device=dev)
    med, best = timed(lambda: ops.hist_nonzero_multi(acts, minmax, 2, 1, B, counts, promotion="nep50"))
    print("multi (27 tensors): median %.1f us  best %.1f us  %.0f GB/s" % (med, best, 4 * n / med / 1e3))
    single = 0.0
    for i, a in enumerate(acts):
        med, best = timed(lambda: ops.hist_nonzero(a, minmax[i, 1:2], B, counts[i], promotion="nep50"), reps=10)
        single += med
        print("layer %2d %-22s %10d elements: %.1f us  %.0f GB/s" % (i, tuple(a.shape), a.numel(), med, 4 * a.numel() / med / 1e3))
    print("27 single-tensor launches: %.1f us in total" % single)


if __name__ == "__main__":
    main()
