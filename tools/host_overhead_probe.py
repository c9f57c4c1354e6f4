"""cProfile of the eager forward of config 1 (host-side cost per fake-quant call)."""
import cProfile
import os
import pstats
import sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantization.mxnet_b200 import model_zoo as Z  # noqa: E402
from quantization.mxnet_b200.quantize import convert  # noqa: E402
from quantization.mxnet_b200.quantize.initialize import qparams_init  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
net = Z.get_model("cifar_resnet20_v1", classes=10).cuda().eval()
convert.convert_model(net, exclude=Z.default_exclusions(net, "cifar_resnet20_v1"))
qparams_init(net)
net.fix_params()
net.quantize_input(True, online=True)
X = torch.randn(128, 3, 32, 32, device="cuda")
with torch.no_grad():
    for _ in range(5):
        net(X)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        net(X)
    pr.disable()
    torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
