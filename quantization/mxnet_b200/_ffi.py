"""ctypes binding of libfq_b200.so (include/fq.h).

Tensors cross the boundary as DLPack ``DLTensor`` structs.  For torch tensors the struct is
filled straight from the tensor (same fields ``__dlpack__`` would export, without the capsule
round trip); any other DLPack producer (e.g. an MXNet NDArray via ``to_dlpack_for_read``) is
accepted through its capsule.  There is deliberately no CPU fallback: a missing library, a CPU
tensor or a failed launch raises.
"""
import ctypes
import os
import threading

import torch

from . import build as _build

_c = ctypes


class FQError(RuntimeError):
    """Raised when a libfq_b200 entry point returns non-zero (cf. mxnet.base.check_call)."""


class DLDevice(_c.Structure):
    _fields_ = [("device_type", _c.c_int32), ("device_id", _c.c_int32)]


class DLDataType(_c.Structure):
    _fields_ = [("code", _c.c_uint8), ("bits", _c.c_uint8), ("lanes", _c.c_uint16)]


class DLTensor(_c.Structure):
    _fields_ = [("data", _c.c_void_p), ("device", DLDevice), ("ndim", _c.c_int32), ("dtype", DLDataType),
                ("shape", _c.POINTER(_c.c_int64)), ("strides", _c.POINTER(_c.c_int64)),
                ("byte_offset", _c.c_uint64)]


P = _c.POINTER(DLTensor)


class FqWeightJob(_c.Structure):
    _fields_ = [("w", P), ("gamma", P), ("beta", P), ("mean", P), ("var", P), ("bias", P), ("rows", _c.c_int64),
                ("bits", _c.c_int32), ("reserved", _c.c_int32), ("w_off", _c.c_int64), ("bias_off", _c.c_int64),
                ("scale_off", _c.c_int64)]

class FqFoldBwdJob(_c.Structure):
    _fields_ = [(n, P) for n in ("dwq", "dbq", "w", "gamma", "mean", "var", "bias", "dw", "dgamma", "dbias", "dbeta")]


kDLCPU, kDLCUDA = 1, 2
_DTYPES = {
    torch.float32: (2, 32), torch.float64: (2, 64),
    torch.int8: (0, 8), torch.int16: (0, 16), torch.int32: (0, 32), torch.int64: (0, 64),
    torch.uint8: (1, 8), torch.uint16: (1, 16),
}

PROMOTION_LEGACY, PROMOTION_NEP50 = 0, 1
STE_IDENTITY, STE_CLIP_MASK = 0, 1
LO_ZERO, LO_NEG_MAX = 0, 1

# name -> (restype, argtypes); must list every symbol of include/fq.h
SIGNATURES = {
    "fq_version": (_c.c_int, []),
    "fq_last_error": (_c.c_char_p, []),
    "fq_workspace_bytes": (_c.c_size_t, []),
    "fq_workspace_init": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "fq_sm_count": (_c.c_int, [_c.POINTER(_c.c_int)]),
    "fq_absmax_rows": (_c.c_int, [P, _c.c_int64, P, _c.c_void_p, _c.c_void_p]),
    "fq_minmax": (_c.c_int, [P, P, _c.c_void_p, _c.c_void_p]),
    "fq_mean_kahan": (_c.c_int, [P, P, _c.c_void_p]),
    "fq_input_range": (_c.c_int, [P, _c.c_int64, P, P, _c.c_void_p, _c.c_void_p]),
    "fq_channel_stats": (_c.c_int, [P, P, P, P, _c.c_void_p, _c.c_void_p]),
    "fq_channel_stats_finish": (_c.c_int, [P, P, P, _c.c_void_p]),
    "fq_scale_from_max": (_c.c_int, [P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, P, _c.c_void_p]),
    "fq_forward_scalar": (_c.c_int, [P, P, P, P, _c.c_void_p]),
    "fq_forward_scalar_host": (_c.c_int, [P, _c.c_float, _c.c_float, _c.c_float, _c.c_float, _c.c_int, P, P,
                                          _c.c_void_p]),
    "fq_forward_rows": (_c.c_int, [P, _c.c_int64, P, P, P, _c.c_void_p]),
    "fq_forward_online": (_c.c_int, [P, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, P, P, P, P, P, P,
                                     _c.c_void_p, _c.c_void_p]),
    "fq_input_plan_create": (_c.c_int, [P, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, P, _c.c_int, P, P, P,
                                        _c.POINTER(_c.c_void_p)]),
    "fq_input_plan_run": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "fq_input_plan_destroy": (_c.c_int, [_c.c_void_p]),
    "fq_forward_from_maxima": (_c.c_int, [P, P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, P, P, P, P, _c.c_void_p]),
    "fq_quant_weight": (_c.c_int, [P, _c.c_int64, _c.c_int, P, P, P, P, P, P, P, P, P, _c.c_void_p, _c.c_void_p]),
    "fq_quant_weight_multi": (_c.c_int, [_c.POINTER(FqWeightJob), _c.c_int, P, P, P, _c.c_void_p, _c.c_void_p]),
    "fq_fold_backward_multi": (_c.c_int, [_c.POINTER(FqFoldBwdJob), _c.c_int, _c.c_void_p]),
    "fq_quant_weight_wino": (_c.c_int, [P, P, P, P, _c.c_int, P, P, _c.c_void_p, _c.c_void_p]),
    "fq_wino_backward": (_c.c_int, [P, P, P, P, P, _c.c_void_p]),
    "fq_ste_backward": (_c.c_int, [P, P, P, P, _c.c_int, _c.c_void_p]),
    "fq_ema_update": (_c.c_int, [P, P, _c.c_double, _c.c_int, _c.c_int, _c.c_void_p]),
    "fq_hist_nonzero": (_c.c_int, [P, P, _c.c_int, _c.c_int, P, P, _c.c_void_p]),
    "fq_hist_nonzero_multi": (_c.c_int, [_c.POINTER(P), _c.c_int, P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, P, P,
                                         _c.c_void_p]),
    "fq_hist_accumulate_f32": (_c.c_int, [P, P, _c.c_int, P, _c.c_void_p]),
    "fq_kl_search": (_c.c_int, [P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, P, P, P, _c.c_void_p]),
    "fq_kl_threshold": (_c.c_int, [P, P, _c.c_int, P, _c.c_void_p]),
    "fq_qconv_pack_input": (_c.c_int, [P, P, _c.c_int, _c.c_int, P, P, _c.c_void_p]),
    "fq_qconv_pack_weight": (_c.c_int, [P, P, P, P, _c.c_void_p]),
    "fq_qconv_igemm": (_c.c_int, [P, P, P, P, P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, P, _c.c_void_p]),
    "fq_nccl_load": (_c.c_int, [_c.c_char_p]),
    "fq_dist_all_reduce": (_c.c_int, [P, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "fq_dist_all_gather": (_c.c_int, [P, P, _c.c_void_p, _c.c_void_p]),
    "fq_dist_input_range": (_c.c_int, [P, P, P, _c.c_void_p, _c.c_void_p]),
    "fq_dist_hist_fold": (_c.c_int, [P, P, _c.c_int, P, _c.c_void_p, _c.c_void_p]),
    "fq_dist_channel_stats": (_c.c_int, [P, P, P, P, _c.c_void_p, _c.c_void_p]),
    "fq_quantize_int8_export": (_c.c_int, [P, P, P, P, _c.c_void_p]),
    "fq_qconv_quantize": (_c.c_int, [P, P, P, P, _c.c_void_p]),
    "fq_qconv_dequantize": (_c.c_int, [P, P, P, P, _c.c_void_p]),
}

_lib = None
_lock = threading.Lock()


def library_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load (building first when the .so is absent and nvcc exists).  Never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            if not build_if_missing:
                raise ImportError("libfq_b200.so is missing: run `python -m quantization.mxnet_b200.build`")
            _build.build()
        lib = _c.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check_call(ret):
    if ret != 0:
        raise FQError(load().fq_last_error().decode("utf-8", "replace"))


class _Arg:
    """A DLTensor plus whatever keeps its memory alive for the duration of the call."""
    __slots__ = ("t", "keep")

    def __init__(self, t, keep):
        self.t = t
        self.keep = keep

    @property
    def ptr(self):
        return _c.byref(self.t)


_PyCapsule_GetPointer = _c.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = _c.c_void_p
_PyCapsule_GetPointer.argtypes = [_c.py_object, _c.c_char_p]


_small_cache = {}      # (data_ptr, shape, dtype, device index) -> _Arg of tiny persistent tensors (qparams, ranges)


def dl(x):
    """Borrow ``x`` as a DLTensor*.  ``None`` -> NULL."""
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            raise FQError("tensor is on %s: quantization.mxnet_b200 has no CPU path" % x.device)
        if not x.is_contiguous():
            raise FQError("tensor must be contiguous (call .contiguous() first)")
        small = x.numel() <= 8
        if small:
            key = (x.data_ptr(), tuple(x.shape), x.dtype, x.device.index)
            hit = _small_cache.get(key)
            if hit is not None:
                return hit
        try:
            code, bits = _DTYPES[x.dtype]
        except KeyError:
            raise FQError("unsupported dtype %s" % x.dtype)
        nd = x.dim()
        shape = (_c.c_int64 * max(nd, 1))(*x.shape)
        t = DLTensor(x.data_ptr(), DLDevice(kDLCUDA, x.device.index or 0), nd, DLDataType(code, bits, 1),
                     shape, None, 0)
        if small:
            # the struct only describes memory; the caller's tensor keeps that memory alive during the call
            arg = _Arg(t, (shape,))
            if len(_small_cache) > 4096:
                _small_cache.clear()
            _small_cache[key] = arg
            return arg
        return _Arg(t, (x, shape))
    if hasattr(x, "__dlpack__") or hasattr(x, "to_dlpack_for_read"):
        cap = x.to_dlpack_for_read() if hasattr(x, "to_dlpack_for_read") else x.__dlpack__()
        ptr = _PyCapsule_GetPointer(cap, b"dltensor")
        t = _c.cast(ptr, P).contents          # DLManagedTensor starts with its DLTensor
        return _Arg(t, (x, cap))
    raise FQError("cannot borrow %r as a DLTensor" % type(x))


def ptr(a):
    return None if a is None else a.ptr


_raw_stream = torch._C._cuda_getCurrentRawStream


def current_stream():
    return _c.c_void_p(_raw_stream(torch.cuda.current_device()))


_workspaces = {}


def workspace_for(dev, raw):
    """Raw address (int) of the workspace of (device index, raw stream handle)."""
    ws = _workspaces.get((dev, raw))
    if ws is None:
        workspace(dev)
        ws = _workspaces[(dev, raw)]
    return ws[2]


def workspace(device=None):
    """Zero-initialised scratch for the fused kernels: one per (device, stream)."""
    if isinstance(device, int):
        dev = device
    else:
        dev = torch.cuda.current_device() if device is None else (torch.device(device).index or 0)
    raw = _raw_stream(dev)
    ws = _workspaces.get((dev, raw))
    if ws is None:
        lib = load()
        nbytes = lib.fq_workspace_bytes()
        with torch.cuda.device(dev):
            ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            check_call(lib.fq_workspace_init(_c.c_void_p(ws.data_ptr()), nbytes, _c.c_void_p(raw)))
        _workspaces[(dev, raw)] = ws = (ws, _c.c_void_p(ws.data_ptr()), ws.data_ptr())
    return ws[1]
