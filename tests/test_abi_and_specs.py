"""CPU-only: the C-ABI library loads and exports every symbol include/mxprune.h declares;
mx_specs validation; the product package never routes through the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mxprune.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mxp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mx_quantization_b200 import _lib, build
    build.build_library()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mxprune.h but not exported"
    assert sorted(_lib.exported_symbols()) == syms        # binding covers the whole header
    assert _lib.load().mxp_abi_version() == _lib.ABI_VERSION


def test_specs_accept_dict_and_userdict():
    import collections
    from mx_quantization_b200 import resolve_specs
    from tests.helpers import mx_specs
    assert resolve_specs(mx_specs()) == (32, False)
    assert resolve_specs(mx_specs(16, True)) == (16, True)
    assert resolve_specs(collections.UserDict(mx_specs(16))) == (16, False)   # MxSpecs is a UserDict
    d = mx_specs(); d["scale_bits"] = 0; d["bfloat"] = 0
    assert resolve_specs(d) == (32, False)


@pytest.mark.parametrize("key,val", [("a_elem_format", "int4"), ("a_elem_format", "fp8_e4m3"),
                                     ("block_size", 16), ("round_mx_output", "floor"),
                                     ("shared_exp_method", "none"), ("bfloat", 12), ("fp", 16),
                                     ("scale_bits", 4)])
def test_specs_reject_off_path(key, val):
    from mx_quantization_b200 import resolve_specs
    from tests.helpers import mx_specs
    d = mx_specs(); d[key] = val
    with pytest.raises(ValueError):
        resolve_specs(d)
    with pytest.raises(ValueError):
        resolve_specs(None)


def test_cpu_tensors_fail_loudly():
    import torch
    import mx_quantization_b200 as m
    from tests.helpers import mx_specs
    x = torch.randn(1, 1, 32, 64)
    with pytest.raises(ValueError, match="no CPU fallback"):
        m.quantize_mxint8(x, mx_specs())
    with pytest.raises(ValueError, match="no CPU fallback"):
        m.pruned_attention(x, x, x, mx_specs(), 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mx_quantization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"
