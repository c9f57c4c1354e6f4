"""One launch of each kernel family touched in round 2, for `ncu --set full` (profiles/r2_ncu_*.txt)."""
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from quantization.mxnet_b200 import ops  # noqa: E402
from quantization.mxnet_b200.nn import Conv2D  # noqa: E402

torch.manual_seed(0)
n = 1 << 28
x = torch.randn(n, device="cuda").abs_()
y = torch.empty_like(x)
cur, qp = torch.empty(1, device="cuda"), torch.empty(4, device="cuda")
mx = torch.tensor([4.0], device="cuda")
for _ in range(2):
    # offline range + tracking (offline_track_tiles_kernel), 2^28 elements, 128 samples
    ops.forward_online(x, 8, False, ops.LO_ZERO, input_max=mx, n_samples=128, out=y, cur_max=cur, qparams=qp)
    # one-pass channel statistics, [128, 64, 512, 64]
    ops.channel_stats(x.view(128, 64, 512, 64))
    # latency-bound online path: range kernel (deferred finish) + self-finishing dependent quantiser, 2^20 elements
    xs = x[:1 << 20].view(128, -1)
    ops.forward_online(xs, 8, False, ops.LO_ZERO, out=y[:1 << 20].view(128, -1), cur_max=cur, qparams=qp)
    # fused fold backward of 8 blocks
    jobs = []
    for c in (64, 128, 256, 512, 64, 128, 256, 512):
        w = torch.randn(c, c // 4, 3, 3, device="cuda")
        jobs.append(dict(dwq=torch.randn_like(w), dbq=torch.randn(c, device="cuda"), w=w, gamma=torch.rand(c, device="cuda") + 0.5,
                         mean=torch.randn(c, device="cuda"), var=torch.rand(c, device="cuda") + 0.5, bias=torch.randn(c, device="cuda")))
    ops.fold_backward_multi(jobs)
    # tcgen05 int8 implicit-GEMM convolution: ResNet-like 3x3, 256 -> 256 channels, 56x56, batch 32
    conv = Conv2D(256, 3, 1, 1, in_channels=256, quantized=True, input_dtype="int8", weight_dtype="int8").cuda()
    xi = torch.randn(32, 256, 56, 56, device="cuda")
    with torch.no_grad():
        out = conv(xi)
torch.cuda.synchronize()
# time the tensor-core convolution against the float-code route (cuDNN fp32 on integer-valued floats)
def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
with torch.no_grad():
    t_tc = timeit(lambda: conv(xi))
    in_rng, unsigned, w_rng = conv._tensor_core_ranges(xi)
    xq, s_in = ops.qconv_pack_input(xi, in_rng, 1, 1)
    wq, s_w = conv._weight_codes(w_rng)
    t_mma = timeit(lambda: ops.qconv_igemm(xq, wq, None, s_in, s_w, (1, 1), 1))
    conv.use_tensor_cores = False
    t_ref = timeit(lambda: conv(xi))
    t_cudnn = timeit(lambda: torch.nn.functional.conv2d(xi, conv.weight, None, 1, 1))
flops = 2.0 * 32 * 56 * 56 * 256 * 256 * 9
print("qconv 32x256x56x56 -> 256, 3x3: tensor-core path %.3f ms (igemm kernel alone %.3f ms = %.1f TOP/s int8), "
      "float-code route %.3f ms, plain cuDNN fp32 conv %.3f ms" % (t_tc, t_mma, flops / t_mma / 1e9, t_ref, t_cudnn))
