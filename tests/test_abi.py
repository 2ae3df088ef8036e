"""The C-ABI library loads without a GPU and exports every symbol include/zkmsm.h declares.
No compute is attempted here; without a device the library must refuse loudly, not fall back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def z():
    import zk_toolkit_b200 as z
    if not os.path.exists(z.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return z


def header_symbols():
    src = open(os.path.join(ROOT, "include", "zkmsm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zkmsm_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(z):
    assert header_symbols() == sorted(z.SYMBOLS)


def test_library_exports_every_declared_symbol(z):
    lib = ctypes.CDLL(z.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in z.load().zkmsm_version()


def test_no_cpu_fallback_without_device(z):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(z.ZkmsmError) as e:
        z.Context(0)
    assert e.value.code == -4
    with pytest.raises(z.ZkmsmError):
        z.Polynomial([1, 2]).eval_with_g1_hidings([z.G1Point.g(), z.G1Point.g()])


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under the package may reference it"""
    pkg = os.path.join(ROOT, "zk-toolkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), os.path.join(dirpath, f)
