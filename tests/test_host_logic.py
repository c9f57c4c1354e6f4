"""Host-side mirror of the reference's Python API, checked without launching any kernel."""
import pytest
import torch
from torch import nn

from quantization.mxnet_b200 import model_zoo as Z
from quantization.mxnet_b200.gluon_compat import NameScope, assign_names, collect_params
from quantization.mxnet_b200.quantize import collect_qparams, convert
from quantization.mxnet_b200.quantize.initialize import qparams_init


def small_net():
    sc = NameScope("net0_")
    net = nn.Sequential(sc(nn.Conv2d(3, 8, 3, bias=False)), sc(nn.BatchNorm2d(8)), sc(nn.ReLU()),
                        sc(nn.Conv2d(8, 8, 3, groups=8, bias=False)), sc(nn.BatchNorm2d(8)), sc(nn.ReLU()),
                        sc(nn.AdaptiveAvgPool2d(1)), sc(nn.Flatten()), sc(nn.Linear(8, 4)))
    return net


def test_gluon_style_names_pair_conv_with_batchnorm():
    net = small_net()
    names = [m.name for m in net]
    assert names[:6] == ["net0_conv0", "net0_batchnorm0", "net0_relu0", "net0_conv1", "net0_batchnorm1", "net0_relu1"]
    assert names[-1] == "net0_dense0"
    p = collect_params(net)
    assert "net0_batchnorm1_gamma" in p and "net0_batchnorm1_running_var" in p and "net0_conv0_weight" in p
    assert list(collect_params(net, "net0_batchnorm0_").keys()) == [
        "net0_batchnorm0_gamma", "net0_batchnorm0_beta", "net0_batchnorm0_running_mean", "net0_batchnorm0_running_var"]
    plain = nn.Sequential(nn.Conv2d(1, 1, 1), nn.BatchNorm2d(1))
    assign_names(plain, "x0_")
    assert [m.name for m in plain] == ["x0_conv0", "x0_batchnorm0"]


def test_convert_model_bookkeeping_matches_reference():
    net = small_net()
    ret = convert.convert_model(net, exclude=[net[0]])
    assert ret is None                                      # convert.py:121 returns nothing
    blocks = net.collect_quantized_blocks()
    assert [b.name for b in blocks] == ["net0_conv1", "net0_dense0"]      # apply order, excluded conv missing
    conv, dense = blocks
    assert conv.quantize_args._fields == ("quantize_input", "in_signed", "in_width", "wt_width", "quant_type",
                                          "fake_bn", "wino_quantize")
    assert dense.quantize_args._fields == ("in_signed", "in_width", "wt_width", "quantize_input", "quant_type")
    assert tuple(conv.quantize_args) == (True, False, 8, 8, "layer", False, "none")
    assert conv.fixed_params == -1 and not hasattr(dense, "fixed_params")
    assert conv.enable_quantize and conv.quantize_input and conv.quantize_input_offline is False
    assert tuple(conv.input_max.shape) == (1,) and conv.input_max.requires_grad is False
    assert float(conv.current_input_max) == 0.0
    assert hasattr(conv, "origin_forward") and not hasattr(net[0], "quantize_args")
    for name in ("update_ema", "collect_quantized_blocks", "quantize_input", "enable_quantize", "disable_quantize",
                 "fix_params"):
        assert callable(getattr(net, name))

    net.quantize_input(enable=True, online=False)
    assert all(b.quantize_input and b.quantize_input_offline for b in blocks)
    net.quantize_input(enable=False)
    assert not any(b.quantize_input for b in blocks)
    net.disable_quantize()
    assert not any(b.enable_quantize for b in blocks)
    net.enable_quantize()
    net.fix_params()
    assert conv.fixed_params == 0 and not hasattr(dense, "fixed_params")   # Dense is never fixed (convert.py:117-121)
    assert list(collect_qparams(net).keys()) == ["net0_conv1_input_max", "net0_dense0_input_max"]


def test_convert_fn_is_exact_type_match_and_custom_fn_wins():
    class MyConv(nn.Conv2d):
        pass
    net = nn.Sequential(MyConv(1, 1, 1), nn.Conv2d(1, 1, 1), nn.Conv2d(1, 1, 1))
    assign_names(net)
    marker = []
    convert.convert_model(net, custom_fn={net[2]: lambda m: marker.append(m)})
    assert not hasattr(net[0], "quantize_args")             # subclass: convert_fn.get(type(m)) misses
    assert hasattr(net[1], "quantize_args")
    assert marker == [net[2]] and not hasattr(net[2], "quantize_args")


def test_quantize_input_asserts_when_block_was_converted_without_it():
    net = small_net()
    fn = {nn.Conv2d: convert.gen_conv2d_converter(quantize_input=False), nn.Linear: convert.gen_dense_converter()}
    convert.convert_model(net, convert_fn=fn)
    with pytest.raises(AssertionError):
        net.quantize_input(enable=True)
    net.quantize_input(enable=False)
    assert not hasattr(net[0], "input_max")


def test_dense_group_maps_to_channel_and_wino_choices():
    lin = nn.Linear(4, 4)
    convert.gen_dense_converter(quant_type="group")(lin)
    assert lin.quantize_args.quant_type == "channel"        # convert_dense.py:83-84
    with pytest.raises(AssertionError):
        convert.gen_conv2d_converter(wino_quantize="F99")
    with pytest.raises(AssertionError):
        convert.gen_conv2d_converter()(nn.Linear(2, 2))


def test_qparams_init_fake_bn_adopts_sibling_batchnorm():
    net = small_net()
    with torch.no_grad():
        net[4].weight.fill_(1.5)
        net[4].bias.fill_(-0.25)
        net[4].running_mean.fill_(0.125)
        net[4].running_var.fill_(2.0)
    fn = {nn.Conv2d: convert.gen_conv2d_converter(fake_bn=True), nn.Linear: convert.gen_dense_converter(),
          nn.BatchNorm2d: convert.bypass_bn}
    convert.convert_model(net, exclude=[net[0], net[1]], convert_fn=fn)
    conv = net[3]
    assert conv.bias is None and tuple(conv.gamma.shape) == (8,)
    assert conv.gamma.requires_grad and conv.beta.requires_grad and not conv.running_var.requires_grad
    with torch.no_grad():
        conv.input_max.fill_(3.0)
    qparams_init(net)
    assert float(conv.input_max) == 0.0
    assert torch.all(conv.gamma == 1.5) and torch.all(conv.beta == -0.25)
    assert torch.all(conv.running_mean == 0.125) and torch.all(conv.running_var == 2.0)
    assert conv.bias is not None and torch.all(conv.bias == 0)          # initialize.py:65-70
    x = torch.randn(2, 8, 5, 5)
    assert net[4](x) is x and net[1](x) is not x                        # bypassed vs excluded BatchNorm


@pytest.mark.parametrize("name,n_conv,n_dense", [("cifar_resnet20_v1", 19, 1), ("mobilenet1.0", 26, 1),
                                                 ("mobilenetv2_1.0", 52, 0), ("resnet50_v1", 52, 1)])
def test_model_zoo_quantised_block_counts(name, n_conv, n_dense):
    net = Z.get_model(name, classes=10)
    convert.convert_model(net, exclude=Z.default_exclusions(net, name))
    blocks = net.collect_quantized_blocks()
    assert sum(isinstance(b, nn.Conv2d) for b in blocks) == n_conv
    assert sum(isinstance(b, nn.Linear) for b in blocks) == n_dense
    params = collect_params(net)
    for b in blocks:
        if isinstance(b, nn.Conv2d):
            assert b.name.replace("conv", "batchnorm") + "_gamma" in params, b.name


def test_weight_job_cache_signature_follows_every_switch_and_every_job_tensor():
    """The net-level weight launch is served from a cached job table (convert_conv2d.prequantize_weights); the table
    is only reused while _weights_signature is unchanged.  Everything that decides the job list or that the table
    holds a raw pointer of must therefore show in the signature (in-place optimizer updates keep the
    storage, so they must NOT invalidate; a new bias, a moved tensor or a flipped switch must)."""
    from quantization.mxnet_b200.quantize.convert import convert_conv2d as C
    net = small_net()
    fn = {nn.Conv2d: convert.gen_conv2d_converter(fake_bn=True, quant_type="channel"),
          nn.Linear: convert.gen_dense_converter(), nn.ReLU: None, nn.BatchNorm2d: convert.bypass_bn}
    convert.convert_model(net, convert_fn=fn)
    qparams_init(net)
    blocks = net.collect_quantized_blocks()
    conv = blocks[0]
    s0 = C._weights_signature(blocks)
    assert C._weights_signature(blocks) == s0                               # stable from call to call
    with torch.no_grad():
        conv.weight.mul_(0.5)                                               # optimizer-style update: same storage
        conv.gamma.add_(1.0)
    assert C._weights_signature(blocks) == s0
    for change, undo in (
            (lambda: setattr(conv, "enable_quantize", False), lambda: setattr(conv, "enable_quantize", True)),
            (lambda: setattr(conv, "fixed_params", 0), lambda: setattr(conv, "fixed_params", -1)),
            (lambda: setattr(blocks[-1], "enable_quantize", False), lambda: setattr(blocks[-1], "enable_quantize", True))):
        change()
        assert C._weights_signature(blocks) != s0
        undo()
        assert C._weights_signature(blocks) == s0
    for name in ("weight", "bias", "gamma", "beta", "running_mean", "running_var"):
        p = getattr(conv, name)
        assert p is not None, name                                          # qparams_init gave the conv a zero bias
        old = p.data
        p.data = old.clone()                                                # what net.to() / a re-pack does
        assert C._weights_signature(blocks) != s0, name
        p.data = old
        assert C._weights_signature(blocks) == s0
    conv.quantize_args = conv.quantize_args._replace(wt_width=4)
    assert C._weights_signature(blocks) != s0
    # the job list itself: CPU weights cannot be batched (no CPU path) -> nothing to launch, per-block path decides
    jobs, owners = C._collect_weight_jobs(blocks)
    assert jobs == [] and owners == []


def test_activation_layout_helpers_alias_channels_last_memory_instead_of_copying():
    """ops._act / ops._ew decide how an activation reaches the kernels (pure host logic, no launch): contiguous tensors
    as they are, channels_last tensors as the contiguous NHWC view of the SAME memory (elementwise results then are
    channels_last as well), anything else -- or an elementwise call with a codes output -- through an NCHW copy."""
    from quantization.mxnet_b200 import ops
    x = torch.randn(2, 8, 5, 3)
    assert ops._act(x) is x
    xc = x.contiguous(memory_format=torch.channels_last)
    v = ops._act(xc)
    assert v.is_contiguous() and tuple(v.shape) == (2, 5, 3, 8) and v.data_ptr() == xc.data_ptr()
    xin, out, ov = ops._ew(xc, None)
    assert xin.data_ptr() == xc.data_ptr() and out.is_contiguous(memory_format=torch.channels_last)
    assert tuple(out.shape) == (2, 8, 5, 3) and ov.is_contiguous() and ov.data_ptr() == out.data_ptr()
    xin, out, ov = ops._ew(xc, None, codes_dtype=torch.int8)            # codes are NCHW: so must x and y be
    assert xin.is_contiguous() and tuple(xin.shape) == (2, 8, 5, 3) and xin.data_ptr() != xc.data_ptr() and ov is out
    given = torch.empty(2, 8, 5, 3)                                       # a caller-provided NCHW output forces NCHW
    xin, out, ov = ops._ew(xc, given)
    assert out is given and ov is given and xin.is_contiguous() and xin.data_ptr() != xc.data_ptr()
    sliced = x[:, ::2]                                                    # neither layout: copied, as before
    assert ops._act(sliced).is_contiguous() and ops._act(sliced).data_ptr() != sliced.data_ptr()
    with pytest.raises(Exception, match="float32"):
        ops._act(x.double())
