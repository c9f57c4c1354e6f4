import sys, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from quantization.mxnet_b200 import ops
for shape in ((16, 1024), (96, 16, 1, 1), (144, 1, 3, 3), (320, 960, 1, 1), (64, 64, 3, 3)):
    w = torch.randn(*shape, device="cuda")
    for rows in (shape[0], 1):
        for _ in range(3):
            ops.quant_weight(w, rows, 8)
    x = torch.randn(128, 4096 // 128 * 4, device="cuda").abs_()
    for _ in range(3):
        ops.forward_online(x, 8, False, ops.LO_ZERO)
torch.cuda.synchronize()
