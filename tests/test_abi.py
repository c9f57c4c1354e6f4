"""CPU-side checks of the drop-in boundary: the library loads and exports exactly what
include/fq.h declares.  No compute call is made here (there is no GPU in this container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fq.h")).read()
    return sorted(set(re.findall(r"FQ_API\s+[\w\s\*]+?\b(fq_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    assert len(names) >= 24
    for must in ("fq_absmax_rows", "fq_forward_scalar", "fq_forward_online", "fq_quant_weight", "fq_ste_backward",
                 "fq_ema_update", "fq_hist_nonzero", "fq_kl_search", "fq_last_error"):
        assert must in names


def test_library_builds_loads_and_exports_every_declared_symbol():
    from quantization.mxnet_b200 import _ffi, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), "libfq_b200.so does not export %s" % name
    # the ctypes signature table covers the header one to one
    assert sorted(_ffi.SIGNATURES) == declared_symbols()
    lib = _ffi.load()
    assert lib.fq_version() >= 100
    assert lib.fq_workspace_bytes() > 65536 * 4


def test_no_cpu_fallback():
    import torch
    from quantization.mxnet_b200 import ops
    from quantization.mxnet_b200._ffi import FQError
    with pytest.raises(FQError, match="no CPU path"):
        ops.forward_scalar_host(torch.ones(16), 1.0, 1.0)
    with pytest.raises(FQError, match="no CPU path"):
        ops.absmax_rows(torch.ones(16), 1)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "quantization")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)
