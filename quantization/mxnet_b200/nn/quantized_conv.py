"""nn/quantized_conv.py of the reference: a stand-alone Conv2D that really quantises to integers
(8-bit hard-coded, no zero point) and dequantises the int32 accumulator.

On the hot path: ``quantize`` / ``_quantize`` / ``dequantize`` -- range reduction and
integer codes run in the CUDA kernels.  The integer convolution itself (outside the
north star) runs on the tensor cores when its operands are 8-bit friendly -- symmetric int8 weights,
int8 inputs (or uint8 inputs with a preset [0, max] range), input channels per group a multiple of 16:
``ops.qconv_pack_input`` writes padded NHWC codes, ``ops.qconv_igemm`` is an implicit GEMM on
``tcgen05.mma.kind::i8`` with exact int32 accumulators in tensor memory and the bias / ReLU /
dequantise epilogue fused (csrc/fq_qconv_mma.cu).  Everything else takes the reference's own route:
a framework convolution on integer-valued float tensors (``F.dot`` on float32 casts, :149-153).
The reference's im2col output-size quirk (``(H - kh + 1) // sh``, :49) is not reproduced: both
routes yield the correct number of windows.
"""
import torch
from torch import nn

from .. import ops

__all__ = ['Conv2D']


def _int2tuple(x):
    return (x, ) * 2 if isinstance(x, int) else tuple(x)


def _quantize(x, min_range, max_range):
    """clip -> scale (max/127 if symmetric else (max-min)/255) -> round -> int32 codes.  (:54-61)
    ``min_range`` / ``max_range``: Python numbers or a (2,) device tensor in ``min_range``."""
    if isinstance(min_range, torch.Tensor) and max_range is None:
        rng = min_range
    else:
        rng = torch.tensor([float(min_range), float(max_range)], dtype=torch.float32, device=x.device)
    return ops.qconv_quantize(x, rng)


def quantize(x, out_type='int8'):
    """(:63-72) int8: symmetric around 0 with max |x|; uint8: [min x, max x] without a zero point."""
    if out_type == 'int8':
        mx = ops.absmax_rows(x, 1)
        rng = torch.cat([-mx, mx])
    elif out_type == 'uint8':
        rng = ops.minmax(x)
    else:
        raise ValueError("unknown out type: ", out_type)
    return ops.qconv_quantize(x, rng)


def dequantize(x, scale):
    """(:74-76) float(x) * scale; ``scale`` may be the product tensor or a pair (s_in, s_w)."""
    if isinstance(scale, tuple):
        return ops.qconv_dequantize(x, scale[0], scale[1])
    one = torch.ones(1, dtype=torch.float32, device=x.device)
    return ops.qconv_dequantize(x, scale, one)


_OVERLAP_MIN_ELEMS = 1 << 23        # below this the layer is bound by host time and a second stream only adds to it
_side_streams = {}


def _side_stream(device):
    """One private side stream per device.  Every use starts with ``side.wait_stream(current)`` and ends with
    ``current.wait_stream(side)``, so memory the caching allocator hands out on it is never reused while the main
    stream still reads an earlier tensor from it."""
    key = torch.device(device).index or 0
    s = _side_streams.get(key)
    if s is None:
        s = _side_streams[key] = torch.cuda.Stream(device=device)
    return s


class Conv2D(nn.Module):
    def __init__(self, channels, kernel_size, strides, padding, in_channels, groups=1,
                 activation=None, use_bias=True, quantized=False,
                 input_dtype='float32', weight_dtype='float32',
                 weight_initializer=None, bias_initializer='zero',
                 prefix=None, params=None):
        super(Conv2D, self).__init__()
        self._channels = channels
        self._in_channels = in_channels
        self._groups = groups
        assert in_channels % groups == 0 and channels % groups == 0
        self._kernel_size = _int2tuple(kernel_size)
        self._strides = _int2tuple(strides)
        self._padding = _int2tuple(padding)
        self._quantized = quantized
        self._input_dtype = input_dtype
        self._weight_dtype = weight_dtype
        self._input_range = None
        self._weight_range = None
        self.use_tensor_cores = True       # set to False to force the framework convolution on float codes
        self.cache_weight_codes = False    # True: keep the int8 weight codes while the weight is unchanged
        self.overlap_weight_prep = True    # large inputs: weight codes are prepared on a side stream meanwhile

        self.weight = nn.Parameter(torch.empty(channels, in_channels // groups, *self._kernel_size))
        nn.init.uniform_(self.weight, -0.07, 0.07)          # mxnet's default Uniform(0.07)
        self.bias = nn.Parameter(torch.zeros(channels)) if use_bias else None
        self.act = nn.ReLU() if activation == 'relu' else None
        if activation not in (None, 'relu'):
            raise NotImplementedError("activation %r" % (activation,))

    def _tensor_core_ranges(self, inputs):
        """(input range2, unsigned codes?, weight range2) when the tcgen05 path applies, else None."""
        if not (self.use_tensor_cores and inputs.is_cuda and (self._in_channels // self._groups) % 16 == 0):
            return None
        dev = inputs.device
        if self._weight_range is None:
            if self._weight_dtype != 'int8':
                return None
            w_rng = None                          # max |w|, taken when the weight codes are (re)built
        else:
            lo, hi = (float(v) for v in self._weight_range)
            if hi != -lo:
                return None
            w_rng = (lo, hi)
        if self._input_range is None:
            if self._input_dtype != 'int8':
                return None             # an automatic uint8 range has no zero point: its codes need not fit 8 bits
            mx = ops.absmax_rows(inputs, 1)       # zero padding does not change max |x|
            return torch.cat([-mx, mx]), False, w_rng
        lo, hi = (float(v) for v in self._input_range)
        if hi == -lo:
            return torch.tensor([lo, hi], dtype=torch.float32, device=dev), False, w_rng
        if lo == 0.0 and hi > 0.0:
            return torch.tensor([lo, hi], dtype=torch.float32, device=dev), True, w_rng       # codes in [0, 255]
        return None

    def _weight_codes(self, w_rng):
        """K-major int8 weight codes + scale for the tensor-core route.  Rebuilt on every forward, as the reference
        quantises its weight on every forward (:118-121); with ``cache_weight_codes = True`` (inference) only when the
        weight's address or autograd version changed -- which in-place edits through ``weight.data`` do not show."""
        w = self.weight
        key = (w.data_ptr(), w._version, w_rng, w.device)
        hit = self.__dict__.get("_fq_wcodes") if self.cache_weight_codes else None
        if hit is None or hit[0] != key:
            wd = w.detach()
            if w_rng is None:
                mx = ops.absmax_rows(wd, 1)
                rng = torch.cat([-mx, mx])
            else:
                rng = torch.tensor(w_rng, dtype=torch.float32, device=w.device)
            hit = (key, ops.qconv_pack_weight(wd, rng))
            if self.cache_weight_codes:
                self.__dict__["_fq_wcodes"] = hit
        return hit[1]

    def forward(self, inputs):
        weight, bias = self.weight, self.bias
        ph, pw = self._padding
        if self._quantized:
            # The weight side (max |w|, int8 K-major codes) does not depend on the activation: for layers whose input
            # side is long enough to hide it, it runs on a side stream beside the input's range and packing passes
            # (fork / join through events, so a CUDA-graph capture records two parallel branches).
            side = None
            if (self.overlap_weight_prep and inputs.is_cuda and inputs.numel() >= _OVERLAP_MIN_ELEMS
                    and self._weight_dtype == 'int8' and self._weight_range is None and not self.cache_weight_codes
                    and self.use_tensor_cores and (self._in_channels // self._groups) % 16 == 0
                    and (self._input_range is not None or self._input_dtype == 'int8')):
                cur = torch.cuda.current_stream(inputs.device)
                side = _side_stream(inputs.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    early = self._weight_codes(None)
            tc = self._tensor_core_ranges(inputs)
            if side is not None:
                cur.wait_stream(side)          # join before anything on this stream touches (or outlives) the codes
            if tc is not None:
                in_rng, unsigned, w_rng = tc
                xq, in_scale = ops.qconv_pack_input(inputs, in_rng, ph, pw, unsigned=unsigned)
                wq, w_scale = early if side is not None else self._weight_codes(w_rng)
                # the float bias goes in as it is: the kernel's epilogue quantises it with b_scale = s_in * s_w,
                # clipped to +- b_scale * 2^31 (:122-127) -- in_scale only exists on the device
                return ops.qconv_igemm(xq, wq, None if bias is None else bias.detach(), in_scale, w_scale, self._strides,
                                       self._groups, relu=self.act is not None)
        inputs = nn.functional.pad(inputs, (pw, pw, ph, ph))
        if self._quantized:
            if self._input_range is None:
                inputs_q, in_scale = quantize(inputs, self._input_dtype)
            else:
                inputs_q, in_scale = _quantize(inputs, *self._input_range)
            if self._weight_range is None:
                weight_q, w_scale = quantize(weight.detach(), self._weight_dtype)
            else:
                weight_q, w_scale = _quantize(weight.detach(), *self._weight_range)
            bias_q = None
            if bias is not None:
                # b_scale = s_in * s_w; clip to +-b_scale * 2^31; round; int32   (:122-127)
                b_scale = in_scale * w_scale
                b_max = b_scale * float(2 ** 31)
                _, bias_q = ops.forward_scalar(bias.detach(), torch.cat([b_scale, b_scale, -b_max, b_max]),
                                               codes_dtype=torch.int32)
            # the reference multiplies float32 casts of the integer codes and casts the result to int32 (:149-153).
            # Its F.dot is exact while the partial sums stay below 2^24.  A framework convolution is not: cuDNN may
            # pick a Winograd / FFT algorithm whose float32 error on sums of ~10^6 exceeds 1 (seen at K = 2304), so
            # this route convolves the codes in float64 -- exact below 2^53 whatever the algorithm -- and rounds
            # before the cast.  (The tensor-core route above is exact by construction and is the fast one.)
            acc = nn.functional.conv2d(inputs_q.double(), weight_q.double(), None, self._strides, 0, 1, self._groups)
            acc = torch.round(acc).to(torch.int32)
            if bias_q is not None:
                acc = acc + bias_q.reshape(1, -1, 1, 1)
            if self.act is not None:
                acc = torch.clamp(acc, min=0)
            return dequantize(acc.contiguous(), (in_scale, w_scale))
        y = nn.functional.conv2d(inputs, weight, bias, self._strides, 0, 1, self._groups)
        return self.act(y) if self.act is not None else y

    def __repr__(self):
        s = '{name}({mapping}, kernel_size={ks}, stride={st}'
        if self._padding != (0,) * len(self._kernel_size):
            s += ', padding={}'.format(self._padding)
        if self._groups != 1:
            s += ', groups={}'.format(self._groups)
        if self.bias is None:
            s += ', bias=False'
        if self.act:
            s += ', {}'.format(self.act)
        s += ')'
        shape = self.weight.shape
        return s.format(name=self.__class__.__name__, ks=self._kernel_size, st=self._strides,
                        mapping='{0} -> {1}'.format(shape[1] if shape[1] else None, shape[0]))
