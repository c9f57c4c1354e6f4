"""Where the e2e step of bench.py goes: H2D copy alone, network forward alone, collect_feature_maps fed from
device-resident and from pinned host images."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from quantization.mxnet_b200.quantize.distribution_calibrate import collect_feature_maps, kl_calibrate_all  # noqa: E402


class Loader:
    def __init__(self, x, n):
        self.x, self.n = x, n

    def __len__(self):
        return self.n

    def __iter__(self):
        for _ in range(self.n):
            yield self.x, None


def wall(fn, reps=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    os.environ.setdefault("TQDM_DISABLE", "1")
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    net = bench.build_net(dev)
    net.disable_quantize()
    X_host = torch.randn(bench.BATCH, 3, 224, 224, generator=torch.Generator().manual_seed(7)).pin_memory()
    X = X_host.to(dev)
    K = 20
    buf = torch.empty_like(X)
    wall(lambda: buf.copy_(X_host, non_blocking=True), 3)
    t = wall(lambda: buf.copy_(X_host, non_blocking=True), 10)
    print("H2D of one batch (%.1f MB, pinned): %.2f ms = %.1f GB/s" % (X_host.numel() * 4 / 1e6, t, X_host.numel() * 4 / t / 1e6))
    with torch.no_grad():
        wall(lambda: net(X), 3)
        print("forward alone (resident input): %.2f ms" % wall(lambda: net(X), K))

    def calib(x, slots):
        hc, mc = collect_feature_maps(net, bench.BINS, Loader(x, K), ctx=dev, ring_slots=slots)
        return kl_calibrate_all(hc, bench.LEVELS, bench.LEVELS, bench.BINS, fm_max=mc)[1].cpu()
    def stats():
        m = torch.cuda.memory_stats()
        return "cudaMallocs %d, reserved %.2f GB" % (m["num_device_alloc"], m["reserved_bytes.all.current"] / 1e9)
    for slots in (1, 32):
        print(stats())
        calib(X, slots)
        print("collect_feature_maps + KL, resident images, ring_slots=%d: %.2f ms/step" % (slots, wall(lambda: calib(X, slots)) / K))
        calib(X_host, slots)
        print("collect_feature_maps + KL, pinned host images, ring_slots=%d: %.2f ms/step" % (slots, wall(lambda: calib(X_host, slots)) / K))


if __name__ == "__main__":
    main()
