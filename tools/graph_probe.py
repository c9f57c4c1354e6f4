"""Per-op time of the online input path under CUDA-graph replay (launch-bound regime): fq_forward_online (range
kernel + quantiser launched as its programmatic dependent) vs. the same two kernels as two plain launches, the
quantiser alone, and the weight path."""
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from quantization.mxnet_b200 import ops  # noqa: E402


def per_op_us(fn, reps=50, inner=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(inner):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * inner)


for lg in (14, 16, 18, 20, 21, 22, 23, 24):
    n = 1 << lg
    x = torch.randn(128, n // 128, device="cuda").abs_()
    y = torch.empty_like(x)
    cur = torch.empty(1, device="cuda")
    qp = torch.empty(4, device="cuda")
    w = torch.randn(max(n // 1024, 1), 1024, device="cuda")
    wq = torch.empty_like(w)

    def fused():
        ops.forward_online(x, 8, False, ops.LO_ZERO, out=y, cur_max=cur, qparams=qp)

    def split():
        ops.forward_online(x, 8, False, ops.LO_ZERO, quantize=False, cur_max=cur, qparams=qp)
        ops.forward_scalar(x, qp, out=y)

    def weights():
        ops.quant_weight(w, w.shape[0], 8, out=wq)

    def plain():
        ops.forward_scalar(x, qp, out=y)
    print("2^%d  fused %.2f us  split %.2f us  plain-quantiser %.2f us  weight-path(coop) %.2f us" % (
        lg, per_op_us(fused), per_op_us(split), per_op_us(plain), per_op_us(weights)), flush=True)
