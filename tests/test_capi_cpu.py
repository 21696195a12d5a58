"""CPU-side checks of the C-ABI library: it builds, loads, exports every declared symbol and
fails loudly (no fallback) when there is no CUDA device.  No compute calls."""
import ctypes as C
import os
import re

import pytest

import teeline_b200 as T
from teeline_b200 import _capi, build as tb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    tb.build()
    return _capi.load()


def test_header_and_exports_agree(lib):
    hdr = open(os.path.join(ROOT, "include", "teeline_cuda.h")).read()
    declared = set(re.findall(r"\b(tl_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"tl_status"}
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in teeline_cuda.h but not exported"


def test_struct_layouts_match_header():
    assert C.sizeof(_capi.Move) == 20  # float, 2 x u32, 2 x u8, u16, u32 k
    assert C.sizeof(_capi.Stats) == 56


def test_version_string(lib):
    assert b"sm_100a" in lib.tl_version()


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(T.TeelineError) as ei:
        T.Context(0)
    assert ei.value.status == 2  # TL_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_package_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "teeline_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "teeline_oracle" not in src, f
