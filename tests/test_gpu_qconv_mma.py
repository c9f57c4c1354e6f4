"""QConv2D's integer convolution on the tensor cores (tcgen05.mma kind::i8, csrc/fq_qconv_mma.cu) against the
reference's own route -- a dot product of float32 casts of the integer codes (nn/quantized_conv.py:149-153), which
tests/test_gpu_api.py checks against the NumPy oracle -- and against an independent float64 integer convolution.
int32 accumulators are exact, so every comparison is bit for bit."""
import numpy as np
import pytest
import torch

from oracle import fq_oracle as O

pytestmark = pytest.mark.gpu
F32 = np.float32

CASES = [
    # N, C, H, W, Cout, k, stride, pad, groups, bias, act, in_dtype / preset range
    dict(n=2, c=16, h=9, w=9, co=8, k=3, s=1, p=1, g=1, bias=True, act=None, inp="int8"),
    dict(n=3, c=32, h=14, w=14, co=40, k=3, s=2, p=1, g=1, bias=True, act="relu", inp="int8"),
    dict(n=2, c=64, h=12, w=10, co=48, k=3, s=1, p=1, g=2, bias=False, act=None, inp="int8"),
    dict(n=2, c=64, h=7, w=7, co=200, k=1, s=1, p=0, g=1, bias=True, act=None, inp="int8"),        # two N tiles
    dict(n=8, c=16, h=28, w=28, co=24, k=3, s=1, p=1, g=1, bias=True, act="relu", inp="int8"),     # 49 M tiles, K=144
    dict(n=2, c=96, h=8, w=8, co=96, k=3, s=1, p=1, g=6, bias=True, act=None, inp="int8"),         # Cg = 16
    dict(n=1, c=256, h=6, w=6, co=128, k=3, s=1, p=0, g=1, bias=False, act=None, inp="int8"),      # K = 2304: 18 k-blocks
    dict(n=4, c=32, h=10, w=10, co=16, k=5, s=2, p=2, g=1, bias=True, act="relu", inp=(0.0, 6.0)),  # uint8 codes
    dict(n=2, c=48, h=9, w=11, co=20, k=(3, 1), s=(1, 2), p=(1, 0), g=1, bias=True, act=None, inp=(-2.5, 2.5)),
    # Cin/groups a multiple of 128: the A operand arrives by TMA in im2col mode (the cases above gather it with cp.async)
    dict(n=3, c=128, h=14, w=14, co=64, k=3, s=2, p=1, g=1, bias=True, act="relu", inp="int8"),    # stride 2, ragged M
    dict(n=2, c=256, h=9, w=11, co=96, k=(3, 1), s=(1, 2), p=(1, 0), g=2, bias=True, act=None, inp="int8"),
    dict(n=5, c=128, h=8, w=8, co=300, k=1, s=1, p=0, g=1, bias=True, act=None, inp=(0.0, 6.0)),   # uint8, two N tiles
    dict(n=8, c=128, h=28, w=28, co=32, k=3, s=1, p=1, g=1, bias=False, act=None, inp="int8"),     # 49 M tiles, tiles span images
    dict(n=1, c=384, h=5, w=7, co=48, k=5, s=1, p=2, g=3, bias=True, act="relu", inp="int8"),      # 5x5 taps, three groups
    # Cin/groups an odd multiple of 64: TMA with 64-byte k-blocks (64 B swizzle, two MMAs per k-block)
    dict(n=2, c=64, h=12, w=10, co=40, k=3, s=1, p=1, g=1, bias=True, act=None, inp="int8"),
    dict(n=2, c=192, h=7, w=9, co=64, k=3, s=2, p=1, g=1, bias=True, act="relu", inp=(0.0, 6.0)),
    dict(n=3, c=128, h=8, w=8, co=272, k=1, s=1, p=0, g=2, bias=False, act=None, inp="int8"),      # Cg = 64, two N tiles per group
]


# how the tensor-core kernel is organised (csrc/fq_qconv_mma.cu): A gathered with cp.async or fetched by TMA in im2col
# mode, one SM per tile or SM pairs (tcgen05 cta_group::2).  The library picks by shape; the tests force each.
MODES = {"auto": {}, "gather": {"FQ_QCONV_TMA_A": "0"}, "tma_1sm": {"FQ_QCONV_2CTA": "0"}, "tma_2sm": {"FQ_QCONV_2CTA": "1"}}


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("case", CASES, ids=lambda c: "c%d_co%d_k%s_g%d_n%d" % (c["c"], c["co"], c["k"], c["g"], c["n"]))
def test_tensor_core_integer_conv_equals_the_float_code_route(case, mode, monkeypatch):
    from quantization.mxnet_b200.nn import Conv2D
    cg = case["c"] // case["g"]
    if (mode == "tma_1sm" and cg % 64 != 0) or (mode == "tma_2sm" and cg % 128 != 0):
        pytest.skip("A by TMA needs Cin/groups % 64 == 0, SM pairs % 128 == 0")
    for k in ("FQ_QCONV_TMA_A", "FQ_QCONV_2CTA"):
        monkeypatch.delenv(k, raising=False)
    for k, v in MODES[mode].items():
        monkeypatch.setenv(k, v)
    torch.manual_seed(case["c"] * 7 + case["co"])
    preset = case["inp"] if isinstance(case["inp"], tuple) else None
    conv = Conv2D(case["co"], case["k"], case["s"], case["p"], in_channels=case["c"], groups=case["g"],
                  activation=case["act"], use_bias=case["bias"], quantized=True,
                  input_dtype="int8" if preset is None else "uint8", weight_dtype="int8").cuda()
    if case["bias"]:
        conv.bias.data.uniform_(-0.2, 0.2)
    x = torch.randn(case["n"], case["c"], case["h"], case["w"], device="cuda")
    if preset is not None:
        conv._input_range = preset
        if preset[0] == 0.0:
            x = x.abs() * 3
    assert conv._tensor_core_ranges(x) is not None
    with torch.no_grad():
        y_tc = conv(x)
        conv.use_tensor_cores = False
        y_ref = conv(x)
    assert y_tc.shape == y_ref.shape

    # the authority: oracle codes, exact integer convolution in float64 on the CPU
    ph, pw = conv._padding
    xp = torch.nn.functional.pad(x, (pw, pw, ph, ph)).cpu().numpy()
    if preset is None:
        xq, s_in = O.qconv_quantize_auto(xp, "int8")
    else:
        xq, s_in = O.qconv_quantize(xp, preset[0], preset[1])
    wq, s_w = O.qconv_quantize_auto(conv.weight.detach().cpu().numpy(), "int8")
    acc = torch.nn.functional.conv2d(torch.from_numpy(xq.astype(np.float64)), torch.from_numpy(wq.astype(np.float64)),
                                     None, conv._strides, 0, 1, case["g"]).numpy().astype(np.int64)
    if case["bias"]:
        bs = F32(s_in * s_w)
        b = conv.bias.detach().cpu().numpy()
        bq = O.roundf((O.clip(b, -bs * F32(2 ** 31), bs * F32(2 ** 31)) / bs).astype(F32)).astype(np.int64)
        acc = acc + bq.reshape(1, -1, 1, 1)
    if case["act"] == "relu":
        acc = np.maximum(acc, 0)
    want = O.qconv_dequantize(acc.astype(np.int32), F32(s_in * s_w))
    assert np.array_equal(y_tc.cpu().numpy().view(np.uint32), want.view(np.uint32))
    # ... and the reference's own route (a product of float casts of the codes, :149-153; here a float64 framework
    # convolution, exact below 2^53 whatever algorithm cuDNN picks) lands on the same integers
    assert torch.equal(y_tc.view(torch.int32), y_ref.view(torch.int32)), (y_tc - y_ref).abs().max().item()
    # deterministic: a second launch gives the same bits (no race between the loader and the tensor core)
    conv.use_tensor_cores = True
    with torch.no_grad():
        assert torch.equal(conv(x).view(torch.int32), y_tc.view(torch.int32))


def test_tensor_core_path_declines_what_it_cannot_represent():
    from quantization.mxnet_b200.nn import Conv2D
    x = torch.rand(2, 3, 8, 8, device="cuda")
    assert Conv2D(8, 3, 1, 1, in_channels=3, quantized=True, input_dtype="int8", weight_dtype="int8").cuda() \
        ._tensor_core_ranges(x) is None                                     # 3 input channels: not a multiple of 16
    x = torch.rand(2, 16, 8, 8, device="cuda")
    conv = Conv2D(8, 3, 1, 1, in_channels=16, quantized=True, input_dtype="uint8", weight_dtype="int8").cuda()
    assert conv._tensor_core_ranges(x) is None                              # automatic uint8 range: codes may not fit
    conv._input_range = (0.5, 4.0)
    assert conv._tensor_core_ranges(x) is None                              # min > 0: codes start above 0 and exceed 255
    with torch.no_grad():
        assert conv(x).shape == (2, 8, 8, 8)                                # ... and the float-code route still runs


def _random_cases(seed, count):
    rng = np.random.RandomState(seed)
    cases = []
    while len(cases) < count:
        cg = int(rng.choice([16, 32, 48, 64, 128, 192, 256]))
        g = int(rng.choice([1, 1, 1, 2, 3]))
        kh, kw = int(rng.choice([1, 2, 3, 5])), int(rng.choice([1, 2, 3, 5]))
        sh, sw = int(rng.choice([1, 1, 2, 3])), int(rng.choice([1, 1, 2, 3]))
        ph, pw = int(rng.randint(0, kh)), int(rng.randint(0, kw))
        h, w = int(rng.randint(kh, 20)), int(rng.randint(kw, 20))
        if h + 2 * ph < kh or w + 2 * pw < kw:
            continue
        n = int(rng.randint(1, 5))
        co_g = int(rng.choice([8, 16, 24, 40, 64, 136, 264]))
        if cg * g * h * w * n > 400000:
            continue
        cases.append(dict(n=n, c=cg * g, h=h, w=w, co=co_g * g, k=(kh, kw), s=(sh, sw), p=(ph, pw), g=g,
                          bias=bool(rng.randint(2)), act="relu" if rng.randint(2) else None,
                          inp="int8" if rng.randint(3) else (0.0, float(rng.uniform(2, 8)))))
    return cases


@pytest.mark.parametrize("mode", list(MODES))
def test_tensor_core_integer_conv_random_shapes(mode, monkeypatch):
    """48 seeded random layers (kernel / stride / padding / groups / channel counts on every loader path, output rows
    shorter than a tile, tiles that wrap rows and images, N tiles that overhang a group) under each kernel
    organisation: bit-exact against a float64 integer convolution of the oracle's codes."""
    from quantization.mxnet_b200.nn import Conv2D
    for k in ("FQ_QCONV_TMA_A", "FQ_QCONV_2CTA"):
        monkeypatch.delenv(k, raising=False)
    for k, v in MODES[mode].items():
        monkeypatch.setenv(k, v)
    ran = 0
    for idx, case in enumerate(_random_cases(20261017, 48)):
        cg = case["c"] // case["g"]
        if (mode == "tma_1sm" and cg % 64 != 0) or (mode == "tma_2sm" and cg % 128 != 0):
            continue
        torch.manual_seed(1000 + idx)
        preset = case["inp"] if isinstance(case["inp"], tuple) else None
        conv = Conv2D(case["co"], case["k"], case["s"], case["p"], in_channels=case["c"], groups=case["g"],
                      activation=case["act"], use_bias=case["bias"], quantized=True,
                      input_dtype="int8" if preset is None else "uint8", weight_dtype="int8").cuda()
        if case["bias"]:
            conv.bias.data.uniform_(-0.2, 0.2)
        x = torch.randn(case["n"], case["c"], case["h"], case["w"], device="cuda")
        if preset is not None:
            conv._input_range = preset
            x = x.abs() * 3
        assert conv._tensor_core_ranges(x) is not None, case
        with torch.no_grad():
            y_tc = conv(x)
        ph, pw = conv._padding
        xp = torch.nn.functional.pad(x, (pw, pw, ph, ph)).cpu().numpy()
        xq, s_in = O.qconv_quantize_auto(xp, "int8") if preset is None else O.qconv_quantize(xp, preset[0], preset[1])
        wq, s_w = O.qconv_quantize_auto(conv.weight.detach().cpu().numpy(), "int8")
        acc = torch.nn.functional.conv2d(torch.from_numpy(xq.astype(np.float64)), torch.from_numpy(wq.astype(np.float64)),
                                         None, conv._strides, 0, 1, case["g"]).numpy().astype(np.int64)
        if case["bias"]:
            bs = F32(s_in * s_w)
            b = conv.bias.detach().cpu().numpy()
            acc = acc + O.roundf((O.clip(b, -bs * F32(2 ** 31), bs * F32(2 ** 31)) / bs).astype(F32)).astype(np.int64).reshape(1, -1, 1, 1)
        if case["act"] == "relu":
            acc = np.maximum(acc, 0)
        want = O.qconv_dequantize(acc.astype(np.int32), F32(s_in * s_w))
        got = y_tc.cpu().numpy()
        assert got.shape == want.shape, (case, got.shape, want.shape)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (mode, idx, case, float(np.abs(got - want).max()))
        ran += 1
    assert ran >= 10


def test_weight_preparation_on_a_side_stream_gives_the_same_layer():
    """Large inputs: the weight side (max |w|, int8 codes) runs on a side stream beside the input's range and packing
    passes.  Same bits as the single-stream order, eagerly and when the fork / join is captured in a CUDA graph."""
    from quantization.mxnet_b200.nn import Conv2D
    from quantization.mxnet_b200.nn import quantized_conv as QC
    torch.manual_seed(5)
    conv = Conv2D(64, 3, 1, 1, in_channels=128, activation="relu", use_bias=True, quantized=True, input_dtype="int8",
                  weight_dtype="int8").cuda()
    conv.bias.data.uniform_(-0.2, 0.2)
    x = torch.randn(8, 128, 96, 96, device="cuda")
    assert x.numel() >= QC._OVERLAP_MIN_ELEMS
    with torch.no_grad():
        conv.overlap_weight_prep = False
        want = conv(x)
        conv.overlap_weight_prep = True
        for _ in range(3):                                   # repeated: the side stream's allocations are recycled
            got = conv(x)
            assert torch.equal(got.view(torch.int32), want.view(torch.int32))
        conv.weight.data.mul_(0.5)                           # new weights are picked up (nothing is cached)
        conv.overlap_weight_prep = False
        want2 = conv(x)
        conv.overlap_weight_prep = True
        assert torch.equal(conv(x).view(torch.int32), want2.view(torch.int32))
        assert not torch.equal(want2, want)
        # captured: two parallel branches in the graph
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                conv(x)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = conv(x)
        x.mul_(1.5)
        ref = None
        g.replay()
        torch.cuda.synchronize()
        conv.overlap_weight_prep = False
        ref = conv(x)
        assert torch.equal(out.view(torch.int32), ref.view(torch.int32))
