"""BASELINE's configurations on their REAL networks and shapes (the config-level parity tests ran on
stand-ins): config 2 on MobileNet-1.0 at 224x224, config 3 on MobileNetV2 with the notebook's converters for one
QAT step, config 4 on ResNet-50 v1 -- batch sizes reduced so that the CPU oracle finishes in seconds; the full
batch sizes run in bench.py and in test_full_size_properties."""
import copy

import numpy as np
import pytest
import torch
from torch import nn

from oracle import build_c as C
from oracle import fq_oracle as O

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.fixture(scope="module")
def Q():
    import types
    from quantization.mxnet_b200 import model_zoo, ops
    from quantization.mxnet_b200.quantize import convert, distribution_calibrate, initialize
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return types.SimpleNamespace(zoo=model_zoo, ops=ops, convert=convert, dc=distribution_calibrate, init=initialize)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def build(Q, name, classes, seed=7, **conv_kwargs):
    torch.manual_seed(seed)
    net = Q.zoo.get_model(name, classes=classes).eval()
    g = torch.Generator().manual_seed(seed + 1)
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):       # non-trivial statistics so that folding is exercised
            m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(1 + 0.2 * torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(1 + 0.1 * torch.randn(m.num_features, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.num_features, generator=g))
    ref = copy.deepcopy(net)
    net = net.cuda()
    dk = {k: v for k, v in conv_kwargs.items() if k in ("weight_width", "input_signed", "input_width", "quantize_input",
                                                        "quant_type")}
    fn = {nn.Conv2d: Q.convert.gen_conv2d_converter(**conv_kwargs), nn.Linear: Q.convert.gen_dense_converter(**dk),
          nn.ReLU: None, nn.BatchNorm2d: Q.convert.bypass_bn if conv_kwargs.get("fake_bn") else None}
    Q.convert.convert_model(net, exclude=Q.zoo.default_exclusions(net, name), convert_fn=fn)
    Q.init.qparams_init(net)
    return net, ref


def capture(net):
    """Per converted block: input x, quantised input xq, quantised weight wq, bias used, and (fake-BN) the raw
    convolution output y_raw of the EMA pre-hook (convert_conv2d.py:149)."""
    rec = {}
    for b in net.collect_quantized_blocks():
        orig = b.origin_forward

        def wrapped(x, w, bias, _b=b, _orig=orig):
            r = rec.setdefault(_b.name, {})
            out = _orig(x, w, bias)
            if w is _b.weight and getattr(_b.quantize_args, "fake_bn", False) and _b.fixed_params != 1 \
                    and not torch.is_grad_enabled():
                r["y_raw"] = out.detach().cpu().numpy()        # the pre-hook's call: raw weight, under no_grad
            else:
                r["xq"] = x.detach().cpu().numpy()
                r["wq"] = w.detach().cpu().numpy()
                r["bias"] = None if bias is None else bias.detach().cpu().numpy()
            return out
        b.origin_forward = wrapped
        b.register_forward_pre_hook(lambda m, x: rec.setdefault(m.name, {}).update(x=x[0].detach().cpu().numpy()))
    return rec


# ---------------------------------------------------------------------------------------------------------------
# config 2: MobileNet-1.0, 224x224, per-channel int8 weights, KL calibration (2048-bin histograms)
# ---------------------------------------------------------------------------------------------------------------
def test_config2_mobilenet1_0_at_224_kl_calibration(Q):
    Q.ops.set_promotion("nep50")          # the regime the reference's NumPy code runs in here (golden fixtures)
    try:
        net, _ = build(Q, "mobilenet1.0", 1000, quant_type="channel")
        net.disable_quantize()            # simulate_quantization.py:298
        g = torch.Generator().manual_seed(11)
        batches = [torch.randn(8, 3, 224, 224, generator=g) * (1.0 + 0.25 * i) for i in range(2)]
        hist_c, max_c = Q.dc.collect_feature_maps(net, 2048, [(b, None) for b in batches], torch.device("cuda"))
        blocks = net.collect_quantized_blocks()
        assert len(blocks) == 27 and set(hist_c.keys()) == set(blocks)
        acts = {}
        hooks = [b.register_forward_hook(lambda m, x, y: acts.setdefault(m.name, []).append(x[0].cpu().numpy()))
                 for b in blocks]
        with torch.no_grad():
            for b in batches:
                net(b.cuda())
        for h in hooks:
            h.remove()
        assert sum(a[0][0].size for a in acts.values()) == 4_993_536      # the element count bench.py is built on
        best_all, th = Q.dc.kl_calibrate_all(hist_c, 256, 256, 2048, fm_max=max_c)
        for i, b in enumerate(blocks):
            want_h, want_m = O.accumulate_histograms(acts[b.name], 2048, "nep50")
            assert np.array_equal(hist_c[b], want_h), b.name             # counts bit-exact, frozen first-batch max
            assert max_c[b] == want_m, b.name
            want_best = C.kl_calibrate(want_h[:2048], 256, 256, 2048, "nep50")[0]
            assert int(best_all[i]) == want_best, (b.name, int(best_all[i]), want_best)
            assert F32(th[i].item()) == O.kl_threshold(want_best, want_m, 2048), b.name
        for b in (blocks[0], blocks[13], blocks[26]):                    # and the NumPy restatement on a few
            assert Q.dc.kl_calibrate(hist_c[b], 256, 256, 2048) == O.kl_calibrate(hist_c[b], 256, 256, 2048, "nep50")
    finally:
        Q.ops.set_promotion("legacy")


# ---------------------------------------------------------------------------------------------------------------
# config 3: MobileNetV2 (CIFAR, 10 classes), the notebook's converters, one QAT step
# ---------------------------------------------------------------------------------------------------------------
NOTEBOOK = dict(quant_type="channel", fake_bn=True, input_width=4, weight_width=4)     # ipynb cell 6 (:109-112)


def test_config3_mobilenetv2_notebook_converters_forward_tensors(Q):
    net, ref = build(Q, "mobilenetv2_1.0", 10, **NOTEBOOK)
    rec = capture(net)
    ref_blocks = {m.name: m for m in ref.modules() if hasattr(m, "name")}
    blocks = net.collect_quantized_blocks()
    assert len(blocks) == 52 and all(isinstance(b, nn.Conv2d) for b in blocks)
    g = torch.Generator().manual_seed(5)
    X = torch.randn(16, 3, 32, 32, generator=g)
    # phase 1 of the notebook: inputs not quantised, ranges tracked (cell 7: net.quantize_input(enable=False))
    net.quantize_input(enable=False)
    with torch.no_grad():
        net(X.cuda())
    net.update_ema()
    state = {}
    for b in blocks:
        cur = O.input_range(rec[b.name]["x"])[0]
        assert F32(b.current_input_max.item()) == cur, b.name
        assert np.array_equal(bits(rec[b.name]["xq"]), bits(rec[b.name]["x"])), b.name     # untouched input
        state[b.name] = O.ema_scalar(np.zeros(1, F32), np.array([cur], F32), 0.9, "legacy")
        assert np.array_equal(bits(b.input_max.detach().cpu().numpy()), bits(state[b.name])), b.name
    # phase 2: offline 4-bit inputs (cell 15: net.quantize_input(enable=True, online=False)), training-mode step
    prev = {b.name: (b.running_mean.detach().cpu().numpy().copy(), b.running_var.detach().cpu().numpy().copy())
            for b in blocks}
    net.quantize_input(enable=True, online=False)
    out = net(X.cuda())
    net.update_ema()
    out.square().mean().backward()
    for b in blocks:
        r = rec[b.name]
        y, _, cur, _ = O.fake_quant_input(r["x"], 4, False, state[b.name][0], "legacy", "conv")
        assert np.array_equal(bits(r["xq"]), bits(y)), b.name                    # 4-bit offline input codes * scale
        rb, bn = ref_blocks[b.name], ref_blocks[b.name.replace("conv", "batchnorm")]
        w2, b2 = O.fold_bn(rb.weight.detach().numpy(), None if rb.bias is None else rb.bias.detach().numpy(),
                           bn.weight.detach().numpy(), bn.bias.detach().numpy(), prev[b.name][0], prev[b.name][1])
        wq, _, _ = O.fake_quant_weight(w2, 4, "channel")
        assert np.array_equal(bits(r["wq"]), bits(wq)), b.name                   # fold + per-channel 4-bit weights
        assert np.array_equal(bits(r["bias"]), bits(b2)), b.name
        # fake-BN batch statistics of the raw convolution output and their EMA (convert_conv2d.py:150-153, convert.py:75-78)
        want_m, want_v = C.channel_stats(r["y_raw"])
        got_m, got_v = b.current_mean.cpu().numpy(), b.current_var.cpu().numpy()
        mag = np.abs(r["y_raw"].astype(np.float64)).mean(axis=(0, 2, 3))
        assert np.all(np.abs(got_m.astype(np.float64) - want_m) <= np.spacing(np.abs(want_m)) + 2.0 ** -23 * mag), b.name
        ulp = np.abs(got_v.view(np.int32).astype(np.int64) - want_v.view(np.int32).astype(np.int64)).max()
        assert ulp <= 2, (b.name, ulp)
        assert np.array_equal(bits(b.running_mean.detach().cpu().numpy()), bits(O.ema_tensor(prev[b.name][0], got_m, 0.9)))
        assert np.array_equal(bits(b.running_var.detach().cpu().numpy()), bits(O.ema_tensor(prev[b.name][1], got_v, 0.9)))
        assert b.weight.grad is not None and b.gamma.grad is not None and b.beta.grad is not None
        assert torch.isfinite(b.weight.grad).all() and torch.isfinite(b.gamma.grad).all()


def _fold_reference(x, conv, gamma, beta, mean, var, bits_w, bits_in, input_max):
    """The un-fused formula of convert_conv2d.py:47-51 + :56-66 + :70-79 as a plain autograd graph; both
    quantisers are straight-through (ste_func.py:43-44)."""
    sd = torch.sqrt(var + 1e-10)
    w2 = (conv.weight * gamma.reshape(-1, 1, 1, 1)) / sd.reshape(-1, 1, 1, 1)
    bias = conv.bias if conv.bias is not None else torch.zeros_like(gamma)
    b2 = (gamma * (bias - mean)) / sd + beta
    wq = torch.from_numpy(O.fake_quant_weight(w2.detach().cpu().numpy(), bits_w, "channel")[0]).cuda()
    w_ste = w2 + (wq - w2).detach()
    xq = torch.from_numpy(O.fake_quant_input(x.detach().cpu().numpy(), bits_in, False, input_max, "legacy", "conv")[0]).cuda()
    x_ste = x + (xq - x).detach()
    return nn.functional.conv2d(x_ste, w_ste, b2, conv.stride, conv.padding, conv.dilation, conv.groups)


@pytest.mark.parametrize("batched", [False, True])
def test_fake_bn_fold_backward_equals_autograd_of_the_unfused_formula(Q, batched):
    """the fold backward (convert_conv2d.py:218-233 here) was only compared with the repo's other path.
    Two fake-BN blocks (dense 3x3 with bias, depthwise 3x3 without); per-block launches and the net-level
    multi-tensor launch (`batched`) against autograd of the formula as the reference writes it."""
    torch.manual_seed(3)
    c1 = nn.Conv2d(6, 8, 3, padding=1, bias=True)
    c2 = nn.Conv2d(8, 8, 3, padding=1, groups=8, bias=False)
    net = nn.Sequential(c1, nn.ReLU(), c2).cuda()
    refs = [copy.deepcopy(c1).cuda(), copy.deepcopy(c2).cuda()]
    conv_fn = Q.convert.gen_conv2d_converter(**NOTEBOOK)
    Q.convert.convert_model(net, convert_fn={nn.Conv2d: conv_fn, nn.ReLU: None})
    net.batch_weight_paths = batched
    g = torch.Generator().manual_seed(9)
    stats = []
    for m in (c1, c2):
        c = m.out_channels
        if m.bias is None:          # initialize.py:65-70 grows a zero bias for fake-BN convolutions
            m.bias = nn.Parameter(torch.zeros(c, device="cuda"))
        with torch.no_grad():
            m.gamma.copy_(1 + 0.3 * torch.randn(c, generator=g))
            m.beta.copy_(0.2 * torch.randn(c, generator=g))
            m.running_mean.copy_(0.3 * torch.randn(c, generator=g))
            m.running_var.copy_(0.5 + torch.rand(c, generator=g))
            m.input_max.fill_(1.25)
        stats.append([t.detach().clone().requires_grad_(t.requires_grad) for t in (m.gamma, m.beta, m.running_mean, m.running_var)])
    net.quantize_input(enable=True, online=False)
    x = torch.rand(4, 6, 10, 10, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)) * 1.5
    x1 = x.clone().requires_grad_(True)
    out = net(x1)
    out.square().sum().backward()

    x2 = x.clone().requires_grad_(True)
    h = _fold_reference(x2, refs[0], *stats[0], 4, 4, 1.25)
    out2 = _fold_reference(torch.relu(h), refs[1], *stats[1], 4, 4, 1.25)
    out2.square().sum().backward()
    assert torch.equal(out, out2)
    pairs = [(x1.grad, x2.grad), (c1.weight.grad, refs[0].weight.grad), (c1.bias.grad, refs[0].bias.grad),
             (c1.gamma.grad, stats[0][0].grad), (c1.beta.grad, stats[0][1].grad),
             (c2.weight.grad, refs[1].weight.grad), (c2.gamma.grad, stats[1][0].grad), (c2.beta.grad, stats[1][1].grad)]
    for got, want in pairs:
        assert got is not None and want is not None
        torch.testing.assert_close(got, want, rtol=2e-5, atol=1e-6)
    # the grown zero bias of the depthwise block receives the gradient of b' w.r.t. b as well
    sd = torch.sqrt(stats[1][3] + 1e-10)
    torch.testing.assert_close(c2.bias.grad, (out2.detach() * 2).sum(dim=(0, 2, 3)) * stats[1][0].detach() / sd,
                               rtol=2e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------
# config 4: ResNet-50 v1, 224x224, per-group 4-bit weights with fake-BN, EMA calibration
# ---------------------------------------------------------------------------------------------------------------
def test_config4_resnet50_v1_fake_bn_per_group_4bit_ema_step(Q):
    net, ref = build(Q, "resnet50_v1", 1000, weight_width=4, quant_type="group", fake_bn=True)
    rec = capture(net)
    ref_blocks = {m.name: m for m in ref.modules() if hasattr(m, "name")}
    blocks = net.collect_quantized_blocks()
    convs = [b for b in blocks if isinstance(b, nn.Conv2d)]
    assert len(blocks) == 53 and len(convs) == 52
    net.quantize_input(enable=True, online=True)              # simulate_quantization.py:322
    X = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    prev = {b.name: (b.running_mean.detach().cpu().numpy().copy(), b.running_var.detach().cpu().numpy().copy())
            for b in convs}
    with torch.no_grad():
        net(X.cuda())
    net.update_ema()
    for b in blocks:
        r = rec[b.name]
        layer = "dense" if isinstance(b, nn.Linear) else "conv"
        y, _, cur, _ = O.fake_quant_input(r["x"], 8, False, None, "legacy", layer)
        assert np.array_equal(bits(r["xq"]), bits(y)), b.name                        # online uint8 inputs
        assert F32(b.current_input_max.item()) == cur, b.name
        want = O.ema_scalar(np.zeros(1, F32), np.array([cur], F32), 0.9, "legacy")
        assert np.array_equal(bits(b.input_max.detach().cpu().numpy()), bits(want)), b.name
    for b in convs:
        r = rec[b.name]
        rb, bn = ref_blocks[b.name], ref_blocks[b.name.replace("conv", "batchnorm")]
        w2, b2 = O.fold_bn(rb.weight.detach().numpy(), None if rb.bias is None else rb.bias.detach().numpy(),
                           bn.weight.detach().numpy(), bn.bias.detach().numpy(), prev[b.name][0], prev[b.name][1])
        wq, _, _ = O.fake_quant_weight(w2, 4, "group", groups=1)                    # G = 1: per-group == per-layer
        assert np.array_equal(bits(r["wq"]), bits(wq)), b.name
        assert np.array_equal(bits(r["bias"]), bits(b2)), b.name
        want_m, want_v = C.channel_stats(r["y_raw"])
        got_m, got_v = b.current_mean.cpu().numpy(), b.current_var.cpu().numpy()
        mag = np.abs(r["y_raw"].astype(np.float64)).mean(axis=(0, 2, 3))
        assert np.all(np.abs(got_m.astype(np.float64) - want_m) <= np.spacing(np.abs(want_m)) + 2.0 ** -23 * mag), b.name
        ulp = np.abs(got_v.view(np.int32).astype(np.int64) - want_v.view(np.int32).astype(np.int64)).max()
        assert ulp <= 2, (b.name, ulp)
        assert np.array_equal(bits(b.running_mean.detach().cpu().numpy()), bits(O.ema_tensor(prev[b.name][0], got_m, 0.9)))
        assert np.array_equal(bits(b.running_var.detach().cpu().numpy()), bits(O.ema_tensor(prev[b.name][1], got_v, 0.9)))


def test_graph_replay_then_update_ema_keeps_the_same_state_tensors(Q):
    """the packed range arenas used to be built lazily by the first update_ema(), AFTER a CUDA graph had
    captured the old pointers.  Now they exist from convert_model on; the natural loop
    `g = GraphedForward(net, x); for x in loader: g(x); net.update_ema()` tracks every batch."""
    from quantization.mxnet_b200.cuda_graph import GraphedForward
    net, _ = build(Q, "cifar_resnet20_v1", 10, weight_width=4, quant_type="channel", fake_bn=True)
    net.quantize_input(enable=True, online=True)
    blocks = net.collect_quantized_blocks()
    ptrs = [(b.input_max.data.data_ptr(), b.current_input_max.data_ptr()) for b in blocks]
    gen = torch.Generator().manual_seed(4)
    xs = [torch.randn(16, 3, 32, 32, generator=gen).cuda() * (1 + i) for i in range(3)]
    eager = copy.deepcopy(net)
    g = GraphedForward(net, xs[0])
    for b, e in zip(blocks, eager.collect_quantized_blocks()):       # same starting state after the warm-up
        with torch.no_grad():
            e.input_max.copy_(b.input_max)
            if getattr(b, "running_mean", None) is not None:
                e.running_mean.copy_(b.running_mean)
                e.running_var.copy_(b.running_var)
    for x in xs:
        y = g(x).clone()
        net.update_ema()
        with torch.no_grad():
            want = eager(x)
        eager.update_ema()
        assert torch.equal(y, want)
        for b, e in zip(blocks, eager.collect_quantized_blocks()):
            assert torch.equal(b.current_input_max, e.current_input_max), b.name
            assert torch.equal(b.input_max.data, e.input_max.data), b.name
            if getattr(b, "running_mean", None) is not None:
                assert torch.equal(b.running_mean.data, e.running_mean.data), b.name
                assert torch.equal(b.running_var.data, e.running_var.data), b.name
    assert ptrs == [(b.input_max.data.data_ptr(), b.current_input_max.data_ptr()) for b in blocks]
    with pytest.raises(RuntimeError, match="captured CUDA graph"):
        net.cpu()                   # moving the net would re-point what the graph holds
    g.release()
