import sys, torch, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from quantization.mxnet_b200 import ops
from oracle import golden_recipes as R
h = np.stack([R.kl_hist_cases()["relu"]] * 27)
hd = torch.from_numpy(h).cuda()
for _ in range(3):
    best, div = ops.kl_search(hd, 256, 256, 2048, promotion="nep50")
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); best, div = ops.kl_search(hd, 256, 256, 2048, promotion="nep50"); b.record(); b.synchronize()
print("kl 27 layers ms", a.elapsed_time(b), best[:3].tolist())
