"""CPU checks of the C-ABI boundary: the shared library builds/loads, exports every symbol that
include/dxb200.h declares, the ctypes table covers them all, and the product path fails loudly
(no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dxb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dx_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from deepxube_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported_and_bound(lib):
    from deepxube_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in dxb200.h but not exported by libdxb200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_struct_sizes_match_header():
    from deepxube_b200 import _lib
    assert ctypes.sizeof(_lib.dx_config) == 4 * 8 + 8 * 2 + 4 * 8
    assert ctypes.sizeof(_lib.dx_net_desc) == 28          # 7 x int32 (act_depth added for heurq_in networks)
    assert ctypes.sizeof(_lib.dx_inst_status) == 8 * 5 + 4 * 4
    assert ctypes.sizeof(_lib.dx_counters) == 64


def test_integration_doc_struct_matches_ctypes():
    """The dx_config mirror printed in INTEGRATION.md section B has the fields, order and size of the real binding."""
    from deepxube_b200 import _lib
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = doc[doc.index("class dx_config(ctypes.Structure)"):]
    block = block[:block.index("]\n") + 1]
    fields = re.findall(r'\("([a-z_0-9]+)", ctypes\.(c_int32 \* 8|c_int32|c_int64)\)', block)
    ctype = {"c_int32": ctypes.c_int32, "c_int64": ctypes.c_int64, "c_int32 * 8": ctypes.c_int32 * 8}

    class Doc(ctypes.Structure):
        _fields_ = [(n, ctype[t]) for n, t in fields]

    assert [n for n, _ in fields] == [n for n, _ in _lib.dx_config._fields_]
    assert ctypes.sizeof(Doc) == ctypes.sizeof(_lib.dx_config)
    for n, _ in fields:
        assert getattr(Doc, n).offset == getattr(_lib.dx_config, n).offset, n


def test_domain_info_is_host_side(lib):
    from deepxube_b200 import _lib
    assert _lib.domain_info("cube3", 0) == (54, 12, 54, 6)
    assert _lib.domain_info("npuzzle", 4) == (16, 4, 16, 16)
    assert _lib.domain_info("npuzzle", 5) == (25, 4, 25, 25)
    assert _lib.domain_info("grid", 7) == (2, 4, 4, 7)
    assert _lib.domain_info("lightsout", 7) == (49, 49, 49, 1)
    with pytest.raises(_lib.DxError):
        _lib.domain_info("npuzzle", 16)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from deepxube_b200 import _lib
    import numpy as np
    with pytest.raises(_lib.DxError, match="no CUDA device"):
        _lib.Engine("cube3")
    with pytest.raises(_lib.DxError, match="no CUDA device"):
        _lib.expand("cube3", 0, np.zeros((1, 54), np.uint8))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under deepxube_b200/ may import it."""
    pkg = os.path.join(ROOT, "deepxube_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
                assert "/root/reference" not in src, os.path.join(dirpath, f)


def test_integration_doc_lists_every_entry_point():
    """INTEGRATION.md section C maps each C-ABI entry point to the reference interface it stands behind."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in _declared() if n not in doc]
    assert not missing, missing
