"""Drop-in API on the GPU vs. the oracle: the converted networks of BASELINE's configs run the
reference's flows (simulate_quantization.py) with every fake-quant tensor checked bit for bit."""
import copy

import numpy as np
import pytest
import torch
from torch import nn

from oracle import fq_oracle as O

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.fixture(scope="module")
def Q():
    import types
    from quantization.mxnet_b200 import model_zoo, ops
    from quantization.mxnet_b200.quantize import convert, distribution_calibrate, freeze, initialize, utils
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return types.SimpleNamespace(zoo=model_zoo, ops=ops, convert=convert, dc=distribution_calibrate, freeze=freeze,
                                 init=initialize, utils=utils)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def build(Q, name, classes, seed=7, **conv_kwargs):
    torch.manual_seed(seed)
    net = Q.zoo.get_model(name, classes=classes).eval()
    # non-trivial BatchNorm statistics so that folding is exercised
    g = torch.Generator().manual_seed(seed + 1)
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(1 + 0.2 * torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(1 + 0.1 * torch.randn(m.num_features, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.num_features, generator=g))
    ref = copy.deepcopy(net)
    net = net.cuda()
    dense_kwargs = {k: v for k, v in conv_kwargs.items() if k in ("weight_width", "input_signed", "input_width",
                                                                  "quantize_input", "quant_type")}
    fn = {nn.Conv2d: Q.convert.gen_conv2d_converter(**conv_kwargs), nn.Linear: Q.convert.gen_dense_converter(**dense_kwargs),
          nn.ReLU: None, nn.BatchNorm2d: Q.convert.bypass_bn if conv_kwargs.get("fake_bn") else None}
    Q.convert.convert_model(net, exclude=Q.zoo.default_exclusions(net, name), convert_fn=fn)
    Q.init.qparams_init(net)
    return net, ref


def capture(net):
    """Record (input, quantised input, quantised weight, bias) of every converted block."""
    rec = {}
    for b in net.collect_quantized_blocks():
        orig = b.origin_forward

        def wrapped(x, w, bias, _b=b, _orig=orig):
            r = rec.setdefault(_b.name, {})     # (the fake-BN EMA pre-hook calls this too, with raw weights)
            r["xq"] = x.detach().cpu().numpy()
            r["wq"] = w.detach().cpu().numpy()
            r["bias"] = None if bias is None else bias.detach().cpu().numpy()
            return _orig(x, w, bias)
        b.origin_forward = wrapped
        b.register_forward_pre_hook(lambda m, x: rec.setdefault(m.name, {}).update(x=x[0].detach().cpu().numpy()))
    return rec


def test_config1_cifar_resnet20_online_uint8_layerwise_and_logits(Q):
    net, ref = build(Q, "cifar_resnet20_v1", 10)
    rec = capture(net)
    net.fix_params()
    net.quantize_input(enable=True, online=True)          # simulate_quantization.py:346-347
    X = torch.randn(32, 3, 32, 32, generator=torch.Generator().manual_seed(7))     # bench_configs.py runs N=128
    with torch.no_grad():
        logits = net(X.cuda()).cpu().numpy()
    blocks = net.collect_quantized_blocks()
    assert len(blocks) == 20
    ref_blocks = {m.name: m for m in ref.modules() if hasattr(m, "name")}
    for b in blocks:
        r = rec[b.name]
        layer = "dense" if isinstance(b, nn.Linear) else "conv"
        y, _, cur, _ = O.fake_quant_input(r["x"], 8, False, None, "legacy", layer)
        assert np.array_equal(bits(r["xq"]), bits(y)), b.name
        assert F32(b.current_input_max.item()) == cur, b.name
        wq, _, _ = O.fake_quant_weight(ref_blocks[b.name].weight.detach().numpy(), 8, "layer")
        assert np.array_equal(bits(r["wq"]), bits(wq)), b.name
        if isinstance(b, nn.Conv2d):
            assert b.fixed_params == 1                     # weights cached (convert_conv2d.py:101-105)
            assert np.array_equal(bits(b.weight.detach().cpu().numpy()), bits(wq))

    # end-to-end: the same pipeline with the ORACLE doing every fake-quant.  The convolutions are
    # framework calls on both sides (cuDNN here and there), so any difference comes from the path
    # under test; north_star's bound is 1e-5 relative.
    def pre(m, x):
        layer = "dense" if isinstance(m, nn.Linear) else "conv"
        y, _, _, _ = O.fake_quant_input(x[0].cpu().numpy(), 8, False, None, "legacy", layer)
        return (torch.from_numpy(y).to(x[0].device),)
    for b in blocks:
        rb = ref_blocks[b.name]
        rb.weight.data = torch.from_numpy(O.fake_quant_weight(rb.weight.detach().numpy(), 8, "layer")[0])
        rb.register_forward_pre_hook(pre)
    with torch.no_grad():
        want = ref.cuda()(X.cuda()).cpu().numpy()
    err = np.abs(logits - want).max() / np.abs(want).max()
    assert err < 1e-5, err
    # against a CPU convolution the summation order differs by ulps, which can flip a rounding tie in
    # a later layer by one quantisation step: the logits then agree only to about that step
    with torch.no_grad():
        want_cpu = ref.cpu()(X).numpy()
    assert np.abs(logits - want_cpu).max() / np.abs(want_cpu).max() < 2e-2

    # second forward: weights are fixed now, only the input path runs
    with torch.no_grad():
        again = net(X.cuda()).cpu().numpy()
    assert np.array_equal(bits(again), bits(logits))


def test_config4_like_fake_bn_per_group_4bit_ema(Q):
    """resnet-style fake-BN + merge, 4-bit per-group weights, EMA ("naive") calibration: §3.3 flow."""
    net, ref = build(Q, "cifar_resnet20_v1", 10, weight_width=4, quant_type="group", fake_bn=True)
    rec = capture(net)
    ref_blocks = {m.name: m for m in ref.modules() if hasattr(m, "name")}
    blocks = net.collect_quantized_blocks()
    convs = [b for b in blocks if isinstance(b, nn.Conv2d)]
    assert all(b.bias is not None for b in convs)              # initialize.py:65-70
    net.quantize_input(enable=True, online=True)
    state = {b.name: np.zeros(1, F32) for b in blocks}
    g = torch.Generator().manual_seed(3)
    for it in range(3):
        X = torch.randn(32, 3, 32, 32, generator=g)
        prev = {b.name: (b.running_mean.detach().cpu().numpy().copy(), b.running_var.detach().cpu().numpy().copy())
                for b in convs}
        with torch.no_grad():
            net(X.cuda())
        net.update_ema()
        for b in blocks:
            cur = O.input_range(rec[b.name]["x"])[0]
            state[b.name] = O.ema_scalar(state[b.name], np.array([cur], F32), 0.9, "legacy")
            assert np.array_equal(bits(b.input_max.detach().cpu().numpy()), bits(state[b.name])), (it, b.name)
        for b in convs:     # fake-BN running statistics (convert.py:75-78)
            want_m = O.ema_tensor(prev[b.name][0], b.current_mean.cpu().numpy(), 0.9)
            want_v = O.ema_tensor(prev[b.name][1], b.current_var.cpu().numpy(), 0.9)
            assert np.array_equal(bits(b.running_mean.detach().cpu().numpy()), bits(want_m)), b.name
            assert np.array_equal(bits(b.running_var.detach().cpu().numpy()), bits(want_v)), b.name
    for b in convs:
        bn = ref_blocks[b.name.replace("conv", "batchnorm")]
        assert np.array_equal(b.gamma.detach().cpu().numpy(), bn.weight.detach().numpy())
        # weight path of the last forward used the running statistics as they were BEFORE the last update
        rb = ref_blocks[b.name]
        w2, b2 = O.fold_bn(rb.weight.detach().numpy(), None if rb.bias is None else rb.bias.detach().numpy(),
                           bn.weight.detach().numpy(), bn.bias.detach().numpy(), prev[b.name][0], prev[b.name][1])
        wq, _, _ = O.fake_quant_weight(w2, 4, "group", groups=1)
        assert np.array_equal(bits(rec[b.name]["wq"]), bits(wq)), b.name
        assert np.array_equal(bits(rec[b.name]["bias"]), bits(b2)), b.name
    # switch to offline inputs with fixed params, as the example does (:336-338)
    net.fix_params()
    net.quantize_input(enable=True, online=False)
    X = torch.randn(32, 3, 32, 32, generator=g)
    with torch.no_grad():
        net(X.cuda())
    for b in blocks:
        layer = "dense" if isinstance(b, nn.Linear) else "conv"
        y, _, _, _ = O.fake_quant_input(rec[b.name]["x"], 8, False, state[b.name][0], "legacy", layer)
        assert np.array_equal(bits(rec[b.name]["xq"]), bits(y)), b.name
    assert all(b.fixed_params == 1 for b in convs)


def test_fake_bn_fold_and_weight_quant_match_oracle(Q):
    net, ref = build(Q, "cifar_resnet20_v1", 10, weight_width=4, quant_type="channel", fake_bn=True)
    rec = capture(net)
    ref_blocks = {m.name: m for m in ref.modules() if hasattr(m, "name")}
    X = torch.randn(8, 3, 32, 32, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        net(X.cuda())
    for b in net.collect_quantized_blocks():
        if not isinstance(b, nn.Conv2d):
            continue
        rb, bn = ref_blocks[b.name], ref_blocks[b.name.replace("conv", "batchnorm")]
        w2, b2 = O.fold_bn(rb.weight.detach().numpy(), None, bn.weight.detach().numpy(), bn.bias.detach().numpy(),
                           bn.running_mean.numpy(), bn.running_var.numpy())
        wq, _, _ = O.fake_quant_weight(w2, 4, "channel")
        assert np.array_equal(bits(rec[b.name]["wq"]), bits(wq)), b.name
        assert np.array_equal(bits(rec[b.name]["bias"]), bits(b2)), b.name


def test_config2_like_kl_calibration_flow(Q):
    """collect_feature_maps + kl_calibrate on the converted net vs. the oracle on the same activations."""
    Q.ops.set_promotion("nep50")          # the regime the golden hist/KL vectors were produced in
    try:
        net, _ = build(Q, "cifar_resnet20_v1", 10, quant_type="channel")
        net.disable_quantize()
        g = torch.Generator().manual_seed(11)
        batches = [torch.randn(16, 3, 32, 32, generator=g) * (1.0 + 0.3 * i) for i in range(3)]
        loader = [(b, None) for b in batches]
        hist_c, max_c = Q.dc.collect_feature_maps(net, 2048, loader, torch.device("cuda"))
        # gather the same activations with plain hooks
        acts = {}
        hooks = [b.register_forward_hook(lambda m, x, y: acts.setdefault(m.name, []).append(x[0].cpu().numpy()))
                 for b in net.collect_quantized_blocks()]
        with torch.no_grad():
            for b in batches:
                net(b.cuda())
        for h in hooks:
            h.remove()
        blocks = net.collect_quantized_blocks()
        assert set(hist_c.keys()) == set(blocks)
        for b in blocks:
            want_h, want_m = O.accumulate_histograms(acts[b.name], 2048, "nep50")
            assert np.array_equal(hist_c[b], want_h), b.name
            assert max_c[b] == want_m
        for b in blocks[:4] + blocks[-2:]:
            best = Q.dc.kl_calibrate(hist_c[b], 256, 256, 2048)
            assert best == O.kl_calibrate(hist_c[b], 256, 256, 2048, "nep50"), b.name
        best_all, th = Q.dc.kl_calibrate_all(hist_c, 256, 256, 2048, fm_max=max_c)
        for i, b in enumerate(blocks[:4]):
            assert int(best_all[i]) == Q.dc.kl_calibrate(hist_c[b], 256, 256, 2048)
            assert F32(th[i].item()) == O.kl_threshold(int(best_all[i]), max_c[b], 2048)
    finally:
        Q.ops.set_promotion("legacy")


def test_collect_feature_maps_loader_forms_and_ring_sizes_agree(Q):
    """Pinned / pageable host batches (double-buffered prefetch), device-resident batches, a ragged last batch
    and every ring size must give the same histograms bit for bit."""
    net, _ = build(Q, "cifar_resnet20_v1", 10, quant_type="channel")
    net.disable_quantize()
    g = torch.Generator().manual_seed(3)
    sizes = [8, 8, 8, 8, 8, 5]
    batches = [torch.randn(n, 3, 32, 32, generator=g) * (1.0 + 0.2 * i) for i, n in enumerate(sizes)]
    dev = torch.device("cuda")
    ref_h, ref_m = Q.dc.collect_feature_maps(net, 2048, [(b.cuda(), None) for b in batches], dev, ring_slots=1)
    blocks = net.collect_quantized_blocks()
    forms = {"pinned": [(b.pin_memory(), None) for b in batches], "pageable": [(b, None) for b in batches]}
    for name, loader in forms.items():
        for slots in (1, 4, 32):
            h, m = Q.dc.collect_feature_maps(net, 2048, loader, dev, ring_slots=slots)
            for b in blocks:
                assert np.array_equal(bits(h[b]), bits(ref_h[b])), (name, slots, b.name)
                assert m[b] == ref_m[b]
    # device-side state of the collectors feeds kl_calibrate_all without an upload
    best = Q.dc.kl_calibrate_all(ref_h, 256, 256, 2048)
    assert best.shape == (len(blocks),)


def test_collect_feature_maps_rejects_negative_activations(Q):
    net, _ = build(Q, "cifar_resnet20_v1", 10)
    net.disable_quantize()
    # include the block right after the un-activated first BatchNorm: its input has negatives
    first = net.features[2][0].body[0]
    Q.convert.gen_conv2d_converter()(first)
    with pytest.raises(AssertionError, match="Activation should >=0"):
        Q.dc.collect_feature_maps(net, 2048, [(torch.randn(4, 3, 32, 32), None)], torch.device("cuda"))


def test_config3_like_qat_step_gradients(Q):
    """Identity STE: gradients equal those of a torch graph whose fake-quant is y = x + (fq(x) - x).detach()."""
    torch.manual_seed(0)
    conv = nn.Conv2d(8, 16, 3, padding=1).cuda()
    lin = nn.Linear(16, 10).cuda()
    ref_conv, ref_lin = copy.deepcopy(conv), copy.deepcopy(lin)
    Q.convert.gen_conv2d_converter(weight_width=4, input_width=4, quant_type="channel")(conv)
    Q.convert.gen_dense_converter(weight_width=4, input_width=4)(lin)
    for m in (conv, lin):
        m.name = "x"
    x = torch.rand(4, 8, 6, 6, device="cuda", requires_grad=True)
    out = lin(torch.relu(conv(x)).mean(dim=(2, 3)))
    out.square().sum().backward()

    def fq_in(t, layer):
        y = O.fake_quant_input(t.detach().cpu().numpy(), 4, False, None, "legacy", layer)[0]
        return t + (torch.from_numpy(y).cuda() - t).detach()

    def fq_w(w, qt):
        y = O.fake_quant_weight(w.detach().cpu().numpy(), 4, qt)[0]
        return w + (torch.from_numpy(y).cuda() - w).detach()
    x2 = x.detach().clone().requires_grad_(True)
    h = torch.relu(nn.functional.conv2d(fq_in(x2, "conv"), fq_w(ref_conv.weight, "channel"), ref_conv.bias, padding=1)).mean(dim=(2, 3))
    out2 = nn.functional.linear(fq_in(h, "dense"), fq_w(ref_lin.weight, "layer"), ref_lin.bias)
    out2.square().sum().backward()
    assert torch.equal(out, out2)
    # cuDNN's backward kernels may pick different (atomics-based) algorithms per call: compare to fp32 accuracy
    for got, want in ((x.grad, x2.grad), (conv.weight.grad, ref_conv.weight.grad), (lin.weight.grad, ref_lin.weight.grad),
                      (conv.bias.grad, ref_conv.bias.grad)):
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-7)


def test_merge_bn_matches_oracle_and_bypasses_bn(Q):
    net, ref = build(Q, "mobilenet1.0", 10)
    ref_blocks = {m.name: m for m in ref.modules() if hasattr(m, "name")}
    X = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(2)).cuda()
    plain = Q.zoo.get_model("mobilenet1.0", classes=10).eval().cuda()
    plain.load_state_dict({k: v for k, v in ref.state_dict().items()})
    with torch.no_grad():
        before = plain(X)
    Q.freeze.merge_bn(plain)
    for m in plain.modules():
        if isinstance(m, nn.Conv2d):
            rb, bn = ref_blocks[m.name], ref_blocks[m.name.replace("conv", "batchnorm")]
            w2, b2 = O.fold_bn(rb.weight.detach().numpy(), None, bn.weight.detach().numpy(), bn.bias.detach().numpy(),
                               bn.running_mean.numpy(), bn.running_var.numpy())
            assert np.array_equal(bits(m.weight.detach().cpu().numpy()), bits(w2)), m.name
            assert np.array_equal(bits(m.bias.detach().cpu().numpy()), bits(b2)), m.name
    with torch.no_grad():
        after = plain(X)
    # BN eps (1e-5) vs the fold's 1e-10: close, not identical -- what tests/test_merge_bn.py eyeballs
    assert torch.allclose(before, after, rtol=1e-3, atol=1e-4)


def test_linear_quantize_ste_call_shapes(Q):
    from quantization.mxnet_b200.quantize.convert import LinearQuantizeSTE
    x = torch.randn(6, 5, 3, 3, device="cuda")
    xn = x.cpu().numpy()
    s = np.float32(0.05)
    y = LinearQuantizeSTE(s, np.float32(1.0), np.float32(-1.0))(x)
    d64 = np.float32(np.float64(s) + 1e-10)
    assert np.array_equal(bits(y.cpu().numpy()), bits(O.fake_quant_scalar(xn, d64, s, -1.0, 1.0)[0]))
    y = LinearQuantizeSTE(0.05, 1.0)(x)                     # clip_min defaults to 0 (ste_func.py:34)
    assert np.array_equal(bits(y.cpu().numpy()),
                          bits(O.fake_quant_scalar(xn, np.float32(0.05 + 1e-10), np.float32(0.05), 0.0, 1.0)[0]))
    sc = torch.rand(6, 1, 1, 1, device="cuda") * 0.1
    y = LinearQuantizeSTE(sc)(x)
    assert np.array_equal(bits(y.cpu().numpy()), bits(O.fake_quant_rows(xn, 6, sc.cpu().numpy())[0]))
    xg = x.clone().requires_grad_(True)
    LinearQuantizeSTE(0.05, 1.0)(xg).sum().backward()
    assert torch.equal(xg.grad, torch.ones_like(xg))        # identity backward, even where clipped


def test_collect_qparams_and_state_dict(Q):
    net, _ = build(Q, "cifar_resnet20_v1", 10)
    qp = Q.utils.collect_qparams(net)
    assert len(qp) == 20 and all(k.endswith("_input_max") for k in qp)
    sd = net.state_dict()
    assert sum(k.endswith("input_max") for k in sd) == 20
    assert not any("current_input_max" in k for k in sd)


def test_qconv2d_integer_path_matches_oracle_and_tracks_float_conv(Q):
    """tests/test_quantized_conv.py of the reference: int path vs float conv vs simulated conv."""
    from quantization.mxnet_b200.nn import Conv2D as MyConv
    torch.manual_seed(1)
    for groups, use_bias in ((1, True), (2, True), (1, False)):
        x = torch.rand(2, 2, 5, 5, device="cuda")
        conv = MyConv(10, 3, 1, 1, in_channels=2, groups=groups, use_bias=use_bias, input_dtype='uint8',
                      weight_dtype='int8', quantized=True).cuda()
        if use_bias:
            conv.bias.data.uniform_(-0.1, 0.1)
        with torch.no_grad():
            y = conv(x)
        # oracle: same integer pipeline in NumPy + framework conv on the integer-valued floats
        xp = torch.nn.functional.pad(x, (1, 1, 1, 1)).cpu().numpy()
        xq, s_in = O.qconv_quantize_auto(xp, "uint8")
        wq, s_w = O.qconv_quantize_auto(conv.weight.detach().cpu().numpy(), "int8")
        acc = torch.nn.functional.conv2d(torch.from_numpy(xq.astype(np.float32)), torch.from_numpy(wq.astype(np.float32)),
                                         None, 1, 0, 1, groups).to(torch.int32).numpy()
        if use_bias:
            bs = F32(s_in * s_w)
            b = conv.bias.detach().cpu().numpy()
            bq = O.roundf((O.clip(b, -bs * F32(2 ** 31), bs * F32(2 ** 31)) / bs).astype(F32)).astype(np.int32)
            acc = acc + bq.reshape(1, -1, 1, 1)
        want = O.qconv_dequantize(acc, F32(s_in * s_w))
        assert np.array_equal(bits(y.cpu().numpy()), bits(want))
        # and it approximates the float convolution (what the reference's script prints)
        ref = torch.nn.functional.conv2d(x, conv.weight, conv.bias, 1, 1, 1, groups)
        assert (y - ref).abs().max() < 0.02 * ref.abs().max()
        conv._quantized = False
        with torch.no_grad():
            assert torch.allclose(conv(x), ref, atol=1e-6)


def test_cuda_graph_replay_equals_eager_and_updates_state(Q):
    from quantization.mxnet_b200.cuda_graph import GraphedForward
    net, _ = build(Q, "cifar_resnet20_v1", 10)
    net.fix_params()
    net.quantize_input(enable=True, online=True)
    g = torch.Generator().manual_seed(5)
    x1 = torch.randn(32, 3, 32, 32, generator=g).cuda()
    x2 = (torch.randn(32, 3, 32, 32, generator=g) * 2).cuda()
    graphed = GraphedForward(net, x1)
    blocks = net.collect_quantized_blocks()
    for x in (x1, x2, x1):
        got = graphed(x).clone()
        cur_g = torch.cat([b.current_input_max for b in blocks]).clone()
        with torch.no_grad():
            want = net(x)
        cur_e = torch.cat([b.current_input_max for b in blocks])
        assert torch.equal(got, want)
        assert torch.equal(cur_g, cur_e)                   # the replay keeps the per-block ranges current
    net.update_ema()
    assert all(float(b.input_max) > 0 for b in blocks)


def test_activation_converter_and_relu6(Q):
    """convert_act.py: ReLU -> ReLU6 and the optional activation-output quantiser (no epsilon, no STE)."""
    relu = nn.ReLU()
    Q.convert.convert_relu_to_relu6(relu)
    x = torch.randn(4, 3, 5, 5, device="cuda") * 5
    assert torch.equal(relu(x), torch.clamp(x, 0, 6))
    act = nn.ReLU().cuda()
    Q.convert.gen_act_converter(width=4)(act)
    act = act.cuda()
    y = act(x)
    a = torch.relu(x).cpu().numpy()
    cur = O.mean_kahan_f32(a.reshape(4, -1).max(axis=1))
    assert F32(act.current_act_max.item()) == cur
    scale = F32(np.float64(cur) / 15)                      # legacy promotion: numpy.float32 / int -> float64
    want = (O.roundf((O.clip(a, 0, cur) / scale).astype(F32)) * scale).astype(F32)
    assert np.array_equal(bits(y.cpu().numpy()), bits(want))


def test_export_quantized_int8_weights_and_thresholds(Q):
    net, ref = build(Q, "cifar_resnet20_v1", 10)
    for i, b in enumerate(net.collect_quantized_blocks()):
        b.input_max.data.fill_(1.0 + i)
    exp = Q.freeze.export_quantized(net)
    ref_blocks = {m.name: m for m in ref.modules() if hasattr(m, "name")}
    assert len(exp) == 20
    for i, (name, e) in enumerate(exp.items()):
        w = ref_blocks[name].weight.detach().numpy()
        mx = np.abs(w).max()
        q, lo, hi = O.quantize_int8_export(w, -mx, mx)
        assert np.array_equal(e["weight_quantize"].cpu().numpy(), q)
        assert F32(e["weight_max"].item()) == hi and F32(e["weight_min"].item()) == lo
        assert e["max_calib_range"] == 1.0 + i and e["min_calib_range"] == 0.0
    # freeze.quantize_params on a plain name list
    params = {"w": torch.randn(8, 4, device="cuda"), "w_min": torch.tensor([-2.0], device="cuda"),
              "w_max": torch.tensor([2.0], device="cuda"), "b": torch.ones(8, device="cuda")}
    qp = Q.freeze.quantize_params(["w_quantize", "b"], params)
    want, lo, hi = O.quantize_int8_export(params["w"].cpu().numpy(), -2.0, 2.0)
    assert np.array_equal(qp["w_quantize"].cpu().numpy(), want) and float(qp["w_quantize_max"]) == 2.0
    assert qp["b"] is params["b"]


def test_multi_tensor_weight_path_gradients_equal_per_block_path(Q):
    """QAT through convert_model with fake-BN: the batched weight launch and its chain rule must give the same
    forward and the same gradients (weight, gamma, beta, bias) as the per-block autograd Functions."""
    torch.backends.cudnn.deterministic = True
    results = []
    for batched in (True, False):
        net, _ = build(Q, "cifar_resnet20_v1", 10, weight_width=4, quant_type="channel", fake_bn=True)
        net.batch_weight_paths = batched
        net.train()
        x = torch.randn(8, 3, 32, 32, generator=torch.Generator().manual_seed(9)).cuda()
        out = net(x)
        out.square().mean().backward()
        grads = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
        results.append((out.detach().clone(), grads))
    torch.backends.cudnn.deterministic = False
    (o1, g1), (o2, g2) = results
    assert torch.equal(o1, o2)
    assert set(g1) == set(g2) and any(k.endswith("gamma") for k in g1) and any(k.endswith("beta") for k in g1)
    for k in g1:
        torch.testing.assert_close(g1[k], g2[k], rtol=1e-5, atol=1e-8, msg=k)


@pytest.mark.parametrize("fake_bn", [False, True])
def test_winograd_domain_weight_quantisation_through_the_converter(Q, fake_bn):
    """gen_conv2d_converter(quant_type='channel', wino_quantize='F43'): 3x3 kernels are quantised in the Winograd
    domain (convert_conv2d.py:71-83), 1x1 kernels per channel as usual; gradients reach the weights."""
    torch.manual_seed(5)
    net = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 8, 1), nn.ReLU(),
                        nn.Conv2d(8, 4, 3, padding=1, groups=2)).cuda()
    conv = Q.convert.gen_conv2d_converter(quant_type="channel", wino_quantize="F43", quantize_input=False,
                                          fake_bn=fake_bn)
    Q.convert.convert_model(net, convert_fn={nn.Conv2d: conv, nn.ReLU: None})
    if fake_bn:
        g = torch.Generator().manual_seed(1)
        for m in net:
            if isinstance(m, nn.Conv2d):
                m.gamma.data.copy_(1 + 0.1 * torch.randn(m.out_channels, generator=g))
                m.beta.data.copy_(0.1 * torch.randn(m.out_channels, generator=g))
                m.running_mean.data.copy_(0.1 * torch.randn(m.out_channels, generator=g))
                m.running_var.data.copy_(1 + 0.2 * torch.rand(m.out_channels, generator=g))
    seen = {}
    for i, m in enumerate(net):
        if isinstance(m, nn.Conv2d):
            orig = m.origin_forward
            m.origin_forward = (lambda x, w, b, _i=i, _o=orig: (seen.__setitem__(_i, (w.detach().cpu().numpy(),
                                None if b is None else b.detach().cpu().numpy())), _o(x, w, b))[1])
    x = torch.randn(2, 3, 8, 8, device="cuda")
    y = net(x)
    for i, m in enumerate(net):
        if not isinstance(m, nn.Conv2d):
            continue
        w = m.weight.detach().cpu().numpy()
        b = None if m.bias is None else m.bias.detach().cpu().numpy()
        if fake_bn:
            w, b = O.fold_bn(w, b, m.gamma.detach().cpu().numpy(), m.beta.detach().cpu().numpy(),
                             m.running_mean.detach().cpu().numpy(), m.running_var.detach().cpu().numpy())
        if tuple(m.kernel_size) == (3, 3):
            want = O.fake_quant_weight_wino(w, "F43", 8)[0]
        else:
            want = O.fake_quant_weight(w, 8, "channel")[0]
        assert np.array_equal(bits(seen[i][0]), bits(want)), i
        if fake_bn:
            assert np.array_equal(bits(seen[i][1]), bits(b)), i
    y.sum().backward()
    for m in net:
        if isinstance(m, nn.Conv2d):
            assert m.weight.grad is not None and torch.isfinite(m.weight.grad).all() and float(m.weight.grad.abs().sum()) > 0
    # fix_params caches the Winograd-quantised weights like any others (:101-105)
    net.fix_params()
    with torch.no_grad():
        y2 = net(x)
        y3 = net(x)
    assert torch.equal(y2, y3) and all(m.fixed_params == 1 for m in net if isinstance(m, nn.Conv2d))


def _two_conv_net(Q):
    torch.manual_seed(0)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.a = nn.Conv2d(3, 8, 3, padding=1)
            self.b = nn.Conv2d(8, 8, 3, padding=1)
            self.use_b = True

        def forward(self, x):
            y = torch.relu(self.a(x))
            return self.b(y) if self.use_b else y
    net = Net().cuda()
    Q.convert.convert_model(net)
    for i, m in enumerate(net.collect_quantized_blocks()):
        m.name = "conv%d" % i
    Q.init.qparams_init(net)
    net.disable_quantize()
    return net


def test_collect_feature_maps_checks_every_batch_like_the_reference(Q):
    """distribution_calibrate.py:35-36 asserts on EVERY batch; round 1 only looked at batch 0.  The histogram kernel
    now raises a device flag for a negative value or a NaN in any batch."""
    net = _two_conv_net(Q)
    g = torch.Generator().manual_seed(1)
    good = [torch.rand(4, 3, 16, 16, generator=g) for _ in range(3)]
    dev = torch.device("cuda")
    h, m = Q.dc.collect_feature_maps(net, 2048, [(b, None) for b in good], dev)
    assert len(h) == 2
    bad = [b.clone() for b in good]
    bad[2][1, 0, 3, 3] = -0.25                       # a negative value in the LAST batch only
    with pytest.raises(AssertionError, match="Activation should >=0"):
        Q.dc.collect_feature_maps(net, 2048, [(b, None) for b in bad], dev)
    bad = [b.clone() for b in good]
    bad[1][0, 2, 5, 5] = float("nan")                # NaN in the middle batch (np.min would be NaN there)
    with pytest.raises(AssertionError, match="Activation should >=0"):
        Q.dc.collect_feature_maps(net, 2048, [(b, None) for b in bad], dev)
    neg_zero = [b.clone() for b in good]
    neg_zero[1][0, 0, 0, 0] = -0.0                   # -0.0 >= 0 is True in NumPy as well
    h2, _ = Q.dc.collect_feature_maps(net, 2048, [(b, None) for b in neg_zero], dev)
    assert len(h2) == 2


@pytest.mark.parametrize("schedule", [(True, False, True, False), (False, True, True, False)],
                         ids=["from_batch0", "first_seen_in_batch1"])
def test_collect_feature_maps_with_a_block_that_is_not_called_in_every_batch(Q, schedule):
    """a block skipped in some batches used to trip the 2048-vs-2049-bin consistency check, and a block
    never called made kl_calibrate_all raise KeyError.  A block's max is frozen by the first batch IT runs in
    (distribution_calibrate.py:97-102), which need not be batch 0."""
    net = _two_conv_net(Q)
    g = torch.Generator().manual_seed(2)
    batches = [torch.rand(4, 3, 16, 16, generator=g) * 300 for _ in range(4)]     # max >= 256: the 2049th bin appears
    dev = torch.device("cuda")
    # the NET decides per forward (the loader is read one batch ahead by the prefetcher)
    plan = iter(schedule)
    inner = net.forward

    def scheduled(x):
        net.use_b = next(plan)
        return inner(x)
    net.forward = scheduled
    h, m = Q.dc.collect_feature_maps(net, 2048, [(b, None) for b in batches], dev)
    net.forward = inner
    blocks = net.collect_quantized_blocks()
    assert set(h.keys()) == set(blocks)
    ran = [i for i, on in enumerate(schedule) if on]
    acts = []
    net.use_b = True
    hook = blocks[1].register_forward_hook(lambda mod, x, y: acts.append(x[0].cpu().numpy()))
    with torch.no_grad():
        for i in ran:
            net(batches[i].cuda())
    hook.remove()
    want_max = F32(acts[0].max())                    # frozen max of the first batch this block saw
    want = np.zeros(2049, F32)
    for a in acts:                                   # the library's default promotion regime is "legacy"
        c = O.histogram_counts(a, 2048, want_max, "legacy").astype(F32)
        want[:len(c)] += c
    got = h[blocks[1]]
    assert m[blocks[1]] == want_max
    assert np.array_equal(got, want[:len(got)]) and want[len(got):].sum() == 0
    # block a ran in every batch
    acts0 = [b.numpy() for b in batches]
    want0_max = F32(acts0[0].max())
    want0 = np.zeros(2049, F32)
    for a in acts0:
        c = O.histogram_counts(a, 2048, want0_max, "legacy").astype(F32)
        want0[:len(c)] += c
    got0 = h[blocks[0]]
    assert m[blocks[0]] == want0_max and np.array_equal(got0, want0[:len(got0)])
    # a block that is never called: no histogram, and the all-layer search skips it
    net.use_b = False
    h, m = Q.dc.collect_feature_maps(net, 2048, [(b, None) for b in batches[:2]], dev)
    assert set(h.keys()) == {blocks[0]}
    best = Q.dc.kl_calibrate_all(h, 256, 256, 2048)
    assert best.shape == (2,) and int(best[0]) >= 256


@pytest.mark.parametrize("mode", ["online", "offline", "range_only"])
@pytest.mark.parametrize("shape,bits_,signed", [((128, 16, 32, 32), 8, False), ((3, 5, 7, 4), 4, True),
                                                ((64, 96, 56, 56), 8, False)])
def test_input_plan_is_forward_online_with_a_persistent_argument_block(Q, mode, shape, bits_, signed):
    """fq_input_plan_run(plan, x, y) == fq_forward_online(x, ..., y): same kernels, captured arguments."""
    ops = Q.ops
    g = torch.Generator().manual_seed(11)
    lo = ops.LO_NEG_MAX if signed else ops.LO_ZERO
    imax = torch.tensor([1.7], device="cuda") if mode == "offline" else None
    quant = mode != "range_only"
    cm_a, cm_b = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    qp_a, qp_b = torch.zeros(4, device="cuda"), torch.zeros(4, device="cuda")
    ps_a = torch.zeros(shape[0], device="cuda")
    ps_b = torch.zeros(shape[0], device="cuda")
    x0 = torch.randn(*shape, generator=g).cuda()
    plan = ops.InputPlan(x0, bits_, signed, lo, input_max=imax, quantize=quant, cur_max=cm_a,
                         qparams=qp_a if quant else None, per_sample=ps_a)
    for seed in range(3):                       # the plan is reused with new activations at new addresses
        x = (torch.randn(*shape, generator=g) * (seed + 1)).cuda()
        y_a = plan.run(x)
        y_b, _, _ = ops.forward_online(x, bits_, signed, lo, input_max=imax, quantize=quant, cur_max=cm_b,
                                       qparams=qp_b, per_sample=ps_b)
        assert (y_a is None) == (y_b is None) == (not quant)
        if quant:
            assert torch.equal(y_a.view(torch.int32), y_b.view(torch.int32))
            assert torch.equal(qp_a.view(torch.int32), qp_b.view(torch.int32))
        assert torch.equal(cm_a, cm_b) and torch.equal(ps_a, ps_b) and float(cm_a) > 0
    if quant:                                   # a quantising plan without an output is an error, not a crash
        from quantization.mxnet_b200 import _ffi
        raw = _ffi._raw_stream(0)
        with pytest.raises(_ffi.FQError, match="y is required"):
            _ffi.check_call(_ffi.load().fq_input_plan_run(plan.handle, x.data_ptr(), 0, _ffi.workspace_for(0, raw), raw))


def test_eager_host_caches_follow_weight_updates_flag_switches_and_moves(Q):
    """The call plans / cached weight job table / persistent no-grad output buffers must never serve stale data:
    in-place weight updates, enable/disable, offline switch, a new bias, ragged batch shapes and net.to() round trips
    all give what a freshly converted copy of the same network gives."""
    def fresh_like(net):
        twin, _ = build(Q, "cifar_resnet20_v1", 10)
        twin.load_state_dict(net.state_dict())
        twin.batch_weight_paths = False                     # the per-block launches, no net-level cache
        for a, b in zip(net.collect_quantized_blocks(), twin.collect_quantized_blocks()):
            b.enable_quantize, b.quantize_input, b.quantize_input_offline = a.enable_quantize, a.quantize_input, \
                a.quantize_input_offline
        return twin

    net, _ = build(Q, "cifar_resnet20_v1", 10)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(16, 3, 32, 32, generator=g).cuda()
    x_small = torch.randn(5, 3, 32, 32, generator=g).cuda()

    def check(tag):
        twin = fresh_like(net)
        with torch.no_grad():
            for inp in (x, x_small, x):
                a, b = net(inp), twin(inp)
                assert torch.equal(a, b), (tag, (a - b).abs().max().item())
        for p, q in zip(net.collect_quantized_blocks(), twin.collect_quantized_blocks()):
            if p.enable_quantize:       # a disabled block tracks nothing: it keeps whatever it saw last
                assert torch.equal(p.current_input_max, q.current_input_max), tag

    check("first")
    check("second forward reuses every cache")
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 4:
                p.mul_(0.75)                                # what an optimizer step does: same storage, new values
    check("in-place weight update")
    net.update_ema(0.5)
    net.quantize_input(True, online=False)
    check("offline ranges")
    net.quantize_input(True, online=True)
    net.disable_quantize()
    check("disabled")
    net.enable_quantize()
    net.collect_quantized_blocks()[3].enable_quantize = False
    check("one block disabled")
    net.enable_quantize()
    net.cpu()
    net.cuda()                                              # state arenas re-packed at new addresses
    check("after a round trip through the host")
    with torch.enable_grad():                               # autograd forwards allocate fresh outputs in between
        net(x).sum().backward()
    check("after an autograd step")


def test_channels_last_activations_are_read_as_they_lie(Q):
    """A channels_last activation is dense memory in N, H, W, C order.  Elementwise quantisers, per-sample / global
    ranges and histograms do not depend on the order inside a sample, so they take that memory as it is (no NCHW
    copy) and an elementwise result comes back channels_last: same values as for the contiguous tensor."""
    ops = Q.ops
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(6, 24, 9, 7, generator=g) * 2).cuda()
    xc = x.contiguous(memory_format=torch.channels_last)
    assert not xc.is_contiguous() and torch.equal(x, xc)
    assert torch.equal(ops.minmax(x), ops.minmax(xc))
    cur_a, cur_b = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    ps_a, ps_b = torch.zeros(6, device="cuda"), torch.zeros(6, device="cuda")
    ops.input_range(x, cur_max=cur_a, per_sample=ps_a)
    ops.input_range(xc, cur_max=cur_b, per_sample=ps_b)
    assert torch.equal(cur_a, cur_b) and torch.equal(ps_a, ps_b)
    pos, posc = x.abs(), xc.abs()
    mx = ops.minmax(pos)[1:2].clone()
    ca = torch.zeros(2049, dtype=torch.int64, device="cuda")
    cb = torch.zeros(2049, dtype=torch.int64, device="cuda")
    ops.hist_nonzero(pos, mx, 2048, ca)
    ops.hist_nonzero(posc, mx, 2048, cb)
    assert torch.equal(ca, cb) and int(ca.sum()) > 0
    for mode_kw in (dict(), dict(input_max=torch.tensor([1.3], device="cuda"))):
        ya, _, qa = ops.forward_online(x, 8, True, ops.LO_NEG_MAX, **mode_kw)
        yb, _, qb = ops.forward_online(xc, 8, True, ops.LO_NEG_MAX, **mode_kw)
        assert yb.is_contiguous(memory_format=torch.channels_last) and not yb.is_contiguous()
        assert torch.equal(ya, yb) and torch.equal(qa, qb)
        assert torch.equal(ops.forward_scalar(xc, qa), ops.forward_scalar(x, qa))
    # with a codes output the layouts of y and codes must agree: the input is copied to NCHW as before
    yc, codes = ops.forward_scalar(xc, qa, codes_dtype=torch.int8)
    yd, codes_d = ops.forward_scalar(x, qa, codes_dtype=torch.int8)
    assert torch.equal(yc, yd) and torch.equal(codes, codes_d)
    # the block-level call plan
    cm, qp = torch.zeros(1, device="cuda"), torch.zeros(4, device="cuda")
    plan = ops.InputPlan(x, 8, True, ops.LO_NEG_MAX, cur_max=cm, qparams=qp)
    yp = plan.run(xc)
    assert yp.is_contiguous(memory_format=torch.channels_last) and torch.equal(yp, ops.forward_online(x, 8, True, ops.LO_NEG_MAX)[0])
    # a converted network in channels_last: calibration histograms equal those of the NCHW network's activations
    net = _two_conv_net(Q)
    batches = [torch.rand(4, 3, 16, 16, generator=g) for _ in range(2)]
    h_a, m_a = Q.dc.collect_feature_maps(net, 2048, [(b, None) for b in batches], torch.device("cuda"))
    blocks = net.collect_quantized_blocks()
    first = {}
    def remember(mod, xin, y):              # (a hook that returns something would replace the block's output)
        first.setdefault("x", xin[0])
    hook = blocks[0].register_forward_hook(remember)
    net_cl = net.to(memory_format=torch.channels_last)
    h_b, m_b = Q.dc.collect_feature_maps(net_cl, 2048, [(b.contiguous(memory_format=torch.channels_last), None) for b in batches],
                                         torch.device("cuda"))
    hook.remove()
    assert not first["x"].is_contiguous()                    # the hooked input really was channels_last
    assert np.array_equal(h_a[blocks[0]], h_b[blocks[0]]) and m_a[blocks[0]] == m_b[blocks[0]]   # the network input itself
