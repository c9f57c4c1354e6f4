"""Small deterministic run over every kernel family, meant to be executed under compute-sanitizer
(memcheck / racecheck / synccheck): sizes are tiny because the tools slow kernels down 10-100x."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantization.mxnet_b200 import ops  # noqa: E402

r = np.random.RandomState(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
x = dev(np.maximum(r.standard_normal((16, 8, 14, 14)), 0).astype(np.float32))          # L = 1568 (short rows)
xl = dev(np.maximum(r.standard_normal((4, 5000)), 0).astype(np.float32))               # long rows, ragged
w = dev((r.standard_normal((32, 8, 3, 3)) * 0.1).astype(np.float32))
wd = dev((r.standard_normal((48, 1, 3, 3)) * 0.1).astype(np.float32))
for t in (x, xl):
    ops.absmax_rows(t, t.shape[0]); ops.absmax_rows(t, 1); ops.minmax(t); ops.input_range(t)
    y, cur, qp = ops.forward_online(t, 8, False, ops.LO_ZERO)
    ops.forward_online(t, 4, True, ops.LO_NEG_MAX, input_max=torch.tensor([1.5], device="cuda"))
    ops.forward_online(t, 8, False, ops.LO_ZERO, quantize=False)
    ops.forward_scalar(t, qp, codes_dtype=torch.uint8)
    ops.ste_backward(t, t, qp, mode=ops.STE_CLIP_MASK)
    ops.forward_rows(t, torch.full((t.shape[0],), 0.01, device="cuda"))
g = torch.ones(32, device="cuda"); z = torch.zeros(32, device="cuda")
ops.quant_weight(w, 32, 8); ops.quant_weight(w, 1, 4); ops.quant_weight(wd, 48, 8)
ops.quant_weight(w, 32, 4, g, z, z, g, None); ops.quant_weight(w, 1, 0, g, z, z, g, z)
plan = ops.WeightPlan([{"w": w, "rows": 32, "bits": 8}, {"w": wd, "rows": 1, "bits": 4},
                       {"w": w, "rows": 32, "bits": 4, "gamma": g, "beta": z, "mean": z, "var": g, "bias": None}])
ops.quant_weight_multi(plan)
ops.channel_stats(dev(r.standard_normal((4, 6, 7, 7)).astype(np.float32)))
st = torch.zeros(8, device="cuda"); ops.ema_update(st, torch.ones(8, device="cuda"), 0.9, True)
counts = torch.zeros(3, 2049, dtype=torch.int64, device="cuda")
mm = torch.stack([ops.minmax(t) for t in (x, xl, x)])
ops.hist_nonzero(x, mm[0, 1:2].clone(), 2048, counts[0])
ops.hist_nonzero_multi([x, xl, x], mm, 2, 1, 2048, counts)
hist = torch.zeros(3, 2049, device="cuda"); ops.hist_accumulate(counts.view(-1), hist.view(-1), True)
ops.hist_nonzero(xl.view(-1)[1:], mm[1, 1:2].clone(), 2048, counts[1])                 # unaligned view: scalar kernel
ring = torch.ones(4, 3, 2049, dtype=torch.int64, device="cuda")
ops.hist_accumulate(ring.view(-1), hist.view(-1), False)                                # four batches in one launch
from quantization.mxnet_b200.quantize.convert import wino_matrix as wm  # noqa: E402   Winograd-domain weights
for name in ("F23", "F43", "F63"):
    G, GI, GTI = (dev(m) for m in wm.winograd_matrices(name))
    ops.quant_weight_wino(w, G, GI, GTI, 8); ops.quant_weight_wino(wd, G, GI, GTI, 4)
    ops.wino_backward(w, G, GI, GTI)
h = dev(np.floor(1e4 * np.exp(-np.arange(2048) / 300.0)).astype(np.float32))
best, div = ops.kl_search(torch.stack([h, h]), 256, 1900, 2048, promotion="nep50")     # 148 candidates x 2 layers
best2, _ = ops.kl_search(h, 1024, 2000, 2048, promotion="legacy")                      # block-per-candidate kernel
ops.kl_threshold(best, torch.ones(2, device="cuda"), 2048)
ops.quantize_int8_export(w, torch.tensor([-1.0, 1.0], device="cuda"))
c, s = ops.qconv_quantize(x, ops.minmax(x)); ops.qconv_dequantize(c, s, s)
# tensor-core QConv2D route: pack (16 channels per thread), TMA + cp.async producers, tcgen05 MMA, TMEM epilogue;
# ragged M (162 rows), two groups, K tail (144 = 128 + 16), float bias quantised in the epilogue, call plan
from quantization.mxnet_b200.nn import Conv2D  # noqa: E402
for cin, cout, k, g_, n_ in ((16, 8, 3, 1, 2), (64, 48, 3, 2, 2), (32, 300, 1, 1, 3)):
    conv = Conv2D(cout, k, 1, k // 2, in_channels=cin, groups=g_, activation="relu", use_bias=True, quantized=True,
                  input_dtype="int8", weight_dtype="int8").cuda()
    with torch.no_grad():
        conv(dev(r.standard_normal((n_, cin, 9, 9)).astype(np.float32)))
# ... and both operands by TMA (Cin/groups % 128 == 0): one SM per tile, then SM pairs (tcgen05 cta_group::2)
for two_cta in ("0", "1"):
    os.environ["FQ_QCONV_2CTA"] = two_cta
    for cin, cout, k, g_, n_ in ((128, 64, 3, 1, 3), (256, 96, 1, 2, 2)):
        conv = Conv2D(cout, k, 1, k // 2, in_channels=cin, groups=g_, activation="relu", use_bias=True, quantized=True,
                      input_dtype="int8", weight_dtype="int8").cuda()
        with torch.no_grad():
            conv(dev(r.standard_normal((n_, cin, 9, 9)).astype(np.float32)))
os.environ.pop("FQ_QCONV_2CTA")
cm, qp2 = torch.zeros(1, device="cuda"), torch.zeros(4, device="cuda")
ip = ops.InputPlan(x, 8, False, ops.LO_ZERO, cur_max=cm, qparams=qp2)
ip.run(x); ip.run(x.clone())
torch.cuda.synchronize()
print("sanitizer probe done", int(best[0]), int(best2[0]))
