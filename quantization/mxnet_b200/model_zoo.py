"""The four networks BASELINE.json's configs name, restated from their published architectures over
torch.nn with Gluon-compatible block names (conv_i <-> batchnorm_i inside one name scope, which is
what qparams_init / merge_bn rely on).  Random init, synthetic data: gluoncv's pretrained weights and
the datasets are unavailable offline.  These are workloads for the hot path, not part of it: the
convolutions themselves are framework calls (cuDNN)."""
import torch
from torch import nn

from .gluon_compat import NameScope

__all__ = ["get_model", "cifar_resnet20_v1", "mobilenet1_0", "mobilenetv2_1_0", "resnet50_v1"]


def _conv(scope, cin, cout, k, s=1, p=0, groups=1, bias=False):
    return scope(nn.Conv2d(cin, cout, k, s, p, groups=groups, bias=bias))


def _bn(scope, c):
    return scope(nn.BatchNorm2d(c, eps=1e-5, momentum=0.1))


def _relu(scope):
    return scope(nn.ReLU(inplace=False))


class _Relu6(nn.Module):
    def forward(self, x):
        return torch.clamp(x, 0, 6)


# ------------------------------------------------------------------------------------------------
# MobileNet 1.0
# ------------------------------------------------------------------------------------------------
class MobileNet(nn.Module):
    def __init__(self, multiplier=1.0, classes=1000, prefix="mobilenet0_"):
        super().__init__()
        sc = NameScope(prefix)
        self.name = prefix.rstrip("_")
        layers = []

        def add_conv(cin, cout, k=1, s=1, p=0, g=1):
            layers.extend([_conv(sc, cin, cout, k, s, p, g), _bn(sc, cout), _relu(sc)])
        c0 = int(32 * multiplier)
        add_conv(3, c0, 3, 2, 1)
        dw = [int(x * multiplier) for x in [32, 64] + [128] * 2 + [256] * 2 + [512] * 6 + [1024]]
        ch = [int(x * multiplier) for x in [64] + [128] * 2 + [256] * 2 + [512] * 6 + [1024] * 2]
        st = [1, 2] * 3 + [1] * 5 + [2, 1]
        for d, c, s in zip(dw, ch, st):
            add_conv(d, d, 3, s, 1, d)
            add_conv(d, c)
        layers.extend([sc(nn.AdaptiveAvgPool2d(1)), sc(nn.Flatten())])
        self.features = nn.Sequential(*layers)
        self.output = sc(nn.Linear(ch[-1], classes))

    def forward(self, x):
        return self.output(self.features(x))


# ------------------------------------------------------------------------------------------------
# MobileNetV2 1.0
# ------------------------------------------------------------------------------------------------
class _LinearBottleneck(nn.Module):
    def __init__(self, scope, cin, c, t, stride):
        super().__init__()
        self.use_shortcut = stride == 1 and cin == c
        self.out = nn.Sequential(
            _conv(scope, cin, cin * t, 1), _bn(scope, cin * t), _Relu6(),
            _conv(scope, cin * t, cin * t, 3, stride, 1, cin * t), _bn(scope, cin * t), _Relu6(),
            _conv(scope, cin * t, c, 1), _bn(scope, c))

    def forward(self, x):
        out = self.out(x)
        return out + x if self.use_shortcut else out


class MobileNetV2(nn.Module):
    def __init__(self, multiplier=1.0, classes=1000, prefix="mobilenetv20_"):
        super().__init__()
        self.name = prefix.rstrip("_")
        fs = NameScope(prefix + "features_")
        c0 = int(32 * multiplier)
        layers = [_conv(fs, 3, c0, 3, 2, 1), _bn(fs, c0), _Relu6()]
        cin_g = [int(x * multiplier) for x in [32] + [16] + [24] * 2 + [32] * 3 + [64] * 4 + [96] * 3 + [160] * 3]
        c_g = [int(x * multiplier) for x in [16] + [24] * 2 + [32] * 3 + [64] * 4 + [96] * 3 + [160] * 3 + [320]]
        ts = [1] + [6] * 16
        strides = [1, 2] * 2 + [1, 1, 2] + [1] * 6 + [2] + [1] * 3
        for i, (cin, c, t, s) in enumerate(zip(cin_g, c_g, ts, strides)):
            layers.append(_LinearBottleneck(fs.child("linearbottleneck%d_" % i), cin, c, t, s))
        last = int(1280 * multiplier) if multiplier > 1.0 else 1280
        layers.extend([_conv(fs, c_g[-1], last, 1), _bn(fs, last), _Relu6(), fs(nn.AdaptiveAvgPool2d(1))])
        self.features = nn.Sequential(*layers)
        os_ = NameScope(prefix + "output_")
        self.output = nn.Sequential(_conv(os_, last, classes, 1), os_(nn.Flatten(), "flatten"))

    def forward(self, x):
        return self.output(self.features(x))


# ------------------------------------------------------------------------------------------------
# ResNet v1 (ImageNet bottleneck and CIFAR basic block)
# ------------------------------------------------------------------------------------------------
class _BottleneckV1(nn.Module):
    def __init__(self, scope, channels, stride, downsample, cin):
        super().__init__()
        mid = channels // 4
        # gluon's BottleneckV1: the 1x1 convolutions keep their bias, the 3x3 and the shortcut do not
        self.body = nn.Sequential(
            _conv(scope, cin, mid, 1, stride, 0, bias=True), _bn(scope, mid), _relu(scope),
            _conv(scope, mid, mid, 3, 1, 1), _bn(scope, mid), _relu(scope),
            _conv(scope, mid, channels, 1, 1, 0, bias=True), _bn(scope, channels))
        self.downsample = None
        if downsample:
            self.downsample = nn.Sequential(_conv(scope, cin, channels, 1, stride), _bn(scope, channels))

    def forward(self, x):
        residual = x if self.downsample is None else self.downsample(x)
        return torch.relu(self.body(x) + residual)


class ResNetV1(nn.Module):
    def __init__(self, layers=(3, 4, 6, 3), channels=(64, 256, 512, 1024, 2048), classes=1000, prefix="resnetv10_"):
        super().__init__()
        self.name = prefix.rstrip("_")
        sc = NameScope(prefix)
        feats = [_conv(sc, 3, channels[0], 7, 2, 3), _bn(sc, channels[0]), _relu(sc), sc(nn.MaxPool2d(3, 2, 1))]
        for i, n in enumerate(layers):
            stride = 1 if i == 0 else 2
            st = sc.child("stage%d_" % (i + 1))
            blocks = [_BottleneckV1(st, channels[i + 1], stride, channels[i + 1] != channels[i], channels[i])]
            blocks += [_BottleneckV1(st, channels[i + 1], 1, False, channels[i + 1]) for _ in range(n - 1)]
            feats.append(nn.Sequential(*blocks))
        feats.extend([sc(nn.AdaptiveAvgPool2d(1)), sc(nn.Flatten())])
        self.features = nn.Sequential(*feats)
        self.output = sc(nn.Linear(channels[-1], classes))

    def forward(self, x):
        return self.output(self.features(x))


class _CifarBasicBlockV1(nn.Module):
    def __init__(self, scope, channels, stride, downsample, cin):
        super().__init__()
        self.body = nn.Sequential(
            _conv(scope, cin, channels, 3, stride, 1), _bn(scope, channels), _relu(scope),
            _conv(scope, channels, channels, 3, 1, 1), _bn(scope, channels))
        self.downsample = None
        if downsample:
            self.downsample = nn.Sequential(_conv(scope, cin, channels, 1, stride), _bn(scope, channels))

    def forward(self, x):
        residual = x if self.downsample is None else self.downsample(x)
        return torch.relu(self.body(x) + residual)


class CifarResNetV1(nn.Module):
    def __init__(self, num_layers=20, classes=10, prefix="cifarresnetv10_"):
        super().__init__()
        assert (num_layers - 2) % 6 == 0
        n = (num_layers - 2) // 6
        channels = [16, 16, 32, 64]
        self.name = prefix.rstrip("_")
        sc = NameScope(prefix)
        feats = [_conv(sc, 3, channels[0], 3, 1, 1), _bn(sc, channels[0])]     # no ReLU here in gluoncv
        for i in range(3):
            stride = 1 if i == 0 else 2
            st = sc.child("stage%d_" % (i + 1))
            blocks = [_CifarBasicBlockV1(st, channels[i + 1], stride, channels[i + 1] != channels[i], channels[i])]
            blocks += [_CifarBasicBlockV1(st, channels[i + 1], 1, False, channels[i + 1]) for _ in range(n - 1)]
            feats.append(nn.Sequential(*blocks))
        feats.extend([sc(nn.AdaptiveAvgPool2d(1)), sc(nn.Flatten())])
        self.features = nn.Sequential(*feats)
        self.output = sc(nn.Linear(channels[-1], classes))

    def forward(self, x):
        return self.output(self.features(x))


def cifar_resnet20_v1(classes=10):
    return CifarResNetV1(20, classes)


def mobilenet1_0(classes=1000):
    return MobileNet(1.0, classes)


def mobilenetv2_1_0(classes=1000):
    return MobileNetV2(1.0, classes)


def resnet50_v1(classes=1000):
    return ResNetV1(classes=classes)


_MODELS = {"cifar_resnet20_v1": cifar_resnet20_v1, "mobilenet1.0": mobilenet1_0, "mobilenetv2_1.0": mobilenetv2_1_0,
           "resnet50_v1": resnet50_v1}


def get_model(name, **kwargs):
    return _MODELS[name](**kwargs)


def default_exclusions(net, model_name, exclude_first_conv=True):
    """examples/simulate_quantization.py:238-244."""
    exclude = []
    if exclude_first_conv:
        exclude.extend([net.features[0], net.features[1]])
    if model_name.startswith('mobilenetv2_'):
        exclude.append(net.output[0])
    if model_name.startswith('cifar_resnet'):
        exclude.extend([net.features[2][0].body[0], net.features[2][0].body[1]])
    return exclude
