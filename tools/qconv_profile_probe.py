import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantization.mxnet_b200 import ops
from quantization.mxnet_b200.nn import Conv2D
conv = Conv2D(256, 3, 1, 1, in_channels=256, quantized=True, input_dtype="int8", weight_dtype="int8").cuda()
x = torch.randn(32, 256, 56, 56, device="cuda")
with torch.no_grad():
    in_rng, unsigned, w_rng = conv._tensor_core_ranges(x)
    xq, s_in = ops.qconv_pack_input(x, in_rng, 1, 1)
    wq, s_w = conv._weight_codes(w_rng)
    for _ in range(2):
        ops.qconv_igemm(xq, wq, None, s_in, s_w, (1, 1), 1)
        torch.cuda.synchronize()
